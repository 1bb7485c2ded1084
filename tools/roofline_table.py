"""Per-kernel roofline table of one eager TIP-cat step from an `ncu --set full` capture.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/roofline_table.py raw.csv workload.json MEASURED_PEAKS.json > profiles/rNN_kernel_roofline.md

For every tipb:: kernel launch: ncu duration (cold cache, serialised), DRAM bytes read+written (`traffic`), DRAM
GB/s, and -- for the kernels SURVEY.md section 8(d) gives a figure for -- the ALGORITHMIC bytes (A) of that launch and
(A)/duration as a fraction of the measured HBM peak.  (A) counts gathered payload rows even when they are served
from shared memory, so fractions above 1 are possible; `traffic` is what actually crossed the DRAM pins.
workload.json comes from `TIPB_DUMP_WORKLOAD=... python tools/one_step.py`."""
import csv
import json
import re
import sys

raw, wl_path, peaks_path = sys.argv[1:4]
W = json.load(open(wl_path))
peak = float(json.load(open(peaks_path))["hbm_gbs"])
E, S, Ssrc, Sneg, R, N = W["E"], W["S_dst"], W["S_src"], W["S_neg"], W["n_rel"], W["n_drug"]
Epp, Np, B = W["E_pp"] + W["n_prot"], W["n_prot"], 32
words = W["mt_words"]

rows = list(csv.reader(l for l in open(raw) if l.startswith('"')))   # ncu prefixes ==PROF== lines
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


unit = {n: rows[1][i] for n, i in col.items()}


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


seen = {}


def algorithmic(name):
    """(bytes, formula) for the n-th launch of this kernel in the step, or None"""
    k = seen.get(name, 0)
    seen[name] = k + 1
    m = re.search(r"k_seg_aggregate(?:_flat)?<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1)) * (1 if "flat" in name else 4)
        return E * (4 + 4 * f) + S * (4 * f + 4), "E(4+4F)+S(4F+4), F=%d" % f
    m = re.search(r"k_decoder_seg<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)", name)
    if m:
        dim, mode = int(m.group(1)) * 4, int(m.group(2))
        ent, seg = (E, S) if mode == 0 else (2 * E, Sneg)
        return ent * (4 + 4 * dim) + seg * (8 * dim + 12), "entries(4+4d)+S(8d+12), d=%d, %s" % (
            dim, "positives: one listing per edge (mirrored)" if mode == 0 else "negatives: doubled plan")
    m = re.search(r"k_node_aggregate<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1)) * 4
        if "gridsmall" in name:
            return None
        return Epp * (4 + 4 * f) + Np * (4 + 8 * f), "E'pp(4+4F)+Np(4+8F), F=%d" % f
    m = re.search(r"k_rgcn_node_fwd_tiled<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1)) * 4
        return S * (4 * f + 4 * B + 12) + N * 4 * f, "S(4F_in+4B+12)+N*4F, F_in=%d" % f
    m = re.search(r"k_rgcn_node_bwd_tiled<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1)) * 4
        return Ssrc * (4 * f + 8 * B + 12) + N * 8 * f, "S(4F_out+8B+12)+N*8F, F_out=%d" % f
    m = re.search(r"k_rgcn_node_fwd_tc<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1))
        return S * (4 * f + 4 * B + 4) + N * 4 * B * f, "S(4F_in+4B+4)+N*4BF (G written), F_in=%d" % f
    m = re.search(r"k_rgcn_node_bwd_tc<\(?(?:int\))?(\d+)", name)
    if m:
        f = int(m.group(1))
        return Ssrc * (4 * f + 8 * B + 4) + N * 8 * B * f, "S(4F_out+4B att+4B d_att+4)+N*8BF (Y read, Q written), F_out=%d" % f
    m = re.search(r"k_pair_pass<\(?(?:int\))?(\d+)", name)
    if m:
        d = int(m.group(1))
        return 2 * E * (8 + 2 * 4 * d), "2E(8+2*4d), d=%d (pos + neg entries, fwd + gradient fused)" % d
    if "k_grp_place" in name:
        return 2 * E * (16 + 8), "2E(16 read + 8 written)"
    if "k_grp_count" in name:
        return E * 16 + 2 * R * N * 4, "16E + 8RN"
    if "k_materialize_main" in name:
        return E * (4 + 4 + 16), "E(4 accepted + 4 bitmap + 16 out)"
    if "k_accept_count" in name:
        return words * 4, "4 B per MT word"
    if "k_compact" in name:
        return words * 4 + int(words * 0.79) * 8, "4/word + 8/accepted"
    if "k_window_scan" in name:
        return W["sum_l"] * 8 + W["sum_w"] * 8, "8 sum_L + 8 sum_W"
    if "k_mt_generate_chunks" in name:
        return words * 4, "4 B per word written"
    return None


print("| # | kernel | grid x block | ncu us | DRAM MB (read+write) = traffic | DRAM GB/s | %% of %.0f GB/s | algorithmic MB (A) | (A)/t GB/s | (A) frac | formula |" % peak)
print("|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
n = 0
agg = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("tipb::", "")
    us = to_us(get(r, "gpu__time_duration.sum"), unit["gpu__time_duration.sum"])
    rd = to_bytes(get(r, "dram__bytes_read.sum"), unit["dram__bytes_read.sum"])
    wr = to_bytes(get(r, "dram__bytes_write.sum"), unit["dram__bytes_write.sum"])
    grid, block = r[col["launch__grid_size"]], r[col["launch__block_size"]]
    tot += us
    n += 1
    a = algorithmic(name) if "tipb" in name or "k_" in name else None
    dram_gbs = (rd + wr) / us / 1e3 if us > 0 else 0.0
    if a:
        ab, formula = a
        print("| %d | `%s` | %s x %s | %.1f | %.1f + %.1f | %.0f | %.1f%% | %.1f | %.0f | %.2f | %s |" % (
            n, short[:60], grid, block, us, rd / 1e6, wr / 1e6, dram_gbs, 100 * dram_gbs / peak, ab / 1e6, ab / us / 1e3,
            ab / us / 1e3 / peak, formula))
    else:
        print("| %d | `%s` | %s x %s | %.1f | %.1f + %.1f | %.0f | %.1f%% | | | | |" % (
            n, short[:60], grid, block, us, rd / 1e6, wr / 1e6, dram_gbs, 100 * dram_gbs / peak))
print()
print("Sum of ncu durations: %.1f us over %d launches (serialised; the CUDA-graph step overlaps the sampler chain with "
      "the encoder, see the bench line for the step time)." % (tot, n))
