"""Instruction mix / stall samples per opcode from `ncu -i rep --page source --csv --kernel-name regex:X --launch-count 1`.
usage: python tools/sass_mix.py source.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for r in rows[h + 1:]:
    if len(r) <= iWi or not r[iI].isdigit():
        continue
    data.append((r[iS].strip(), int(r[iN]), int(r[iI]), int(r[iW] or 0), int(r[iWi] or 0)))
ts, ti = sum(d[1] for d in data), sum(d[2] for d in data)
print(rows[0][1][:100])
print("warp instructions %d, stall samples %d, shared wavefronts %d (ideal %d)" % (ti, ts, sum(d[3] for d in data), sum(d[4] for d in data)))
op, smp = collections.Counter(), collections.Counter()
for s, n, i, w, wi in data:
    parts = s.split()
    k = parts[1] if parts[0].startswith("@") else parts[0]
    k = k.split(".")[0] + ("." + k.split(".")[1] if k.startswith(("LDS", "LDG", "STG", "STS")) and "." in k else "")
    op[k] += i
    smp[k] += n
for k, v in op.most_common(22):
    print("%-12s instr %10d (%4.1f%%)   samples %6d (%4.1f%%)" % (k, v, 100 * v / ti, smp[k], 100 * smp[k] / ts))
