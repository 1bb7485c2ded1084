"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST step.
usage: python tools/breakdown.py launches.csv n_steps  (the list holds n_steps eager steps; the last one is steady state)"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3, r["Stream"]))
# the step boundary: the fused Adam kernels end a step; take everything after the second-to-last Adam group
idx = [i for i, r in enumerate(rows) if "k_mt_generate" in r[0]]
start = idx[-2] + 1 if len(idx) >= 2 else 0
end = idx[-1] + 1
last = rows[start:end]
agg = collections.OrderedDict()
for name, us, st in last:
    name = re.sub(r"\(.*", "", name)[:90]
    t, n = agg.get(name, (0.0, 0))
    agg[name] = (t + us, n + 1)
tot = sum(t for t, _ in agg.values())
print("step total %.1f us over %d launches (serialised, cold-cache ncu times; compare shares)" % (tot, len(last)))
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%10.1f us %5.1f%%  x%3d  %s" % (t, 100 * t / tot, n, name))
