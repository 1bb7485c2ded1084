"""ncu raw CSV (metrics pass incl. dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum over the
kernels of eager steps) -> profiles/kernel_traffic.json: per kernel (short name) the average DRAM traffic and ncu
duration of one launch.  bench.py reads it for `roofline.traffic` of whichever kernel it finds dominant.
usage: python tools/kernel_traffic.py raw.csv out.json "<source note>" """
import csv
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tools.roofline import short_name  # noqa: E402

raw, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(l for l in open(raw) if l.startswith('"')))   # ncu prefixes ==PROF== lines
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def val(r, name, table):
    i = col[name]
    v = r[i].replace(",", "")
    return float(v) * table.get(units[i], 1) if v not in ("", "n/a") else 0.0


agg = {}
for r in rows[2:]:
    name = short_name(r[col["Kernel Name"]])
    a = agg.setdefault(name, {"launches": 0, "rd": 0.0, "wr": 0.0, "us": 0.0})
    a["launches"] += 1
    a["rd"] += val(r, "dram__bytes_read.sum", scale_b)
    a["wr"] += val(r, "dram__bytes_write.sum", scale_b)
    a["us"] += val(r, "gpu__time_duration.sum", scale_t)
kernels = {k: {"launches_in_capture": a["launches"], "dram_read_bytes": int(a["rd"] / a["launches"]),
               "dram_write_bytes": int(a["wr"] / a["launches"]), "traffic_bytes": int((a["rd"] + a["wr"]) / a["launches"]),
               "ncu_duration_us": round(a["us"] / a["launches"], 2)} for k, a in agg.items()}
# ncu prints a bool template argument as 1 / 0, CUPTI (bench.py) as true / false: alias those names
for k in list(kernels):
    for a, b in (("<1>", "<true>"), ("<0>", "<false>")):
        if k.endswith(a) and k[:-len(a)] + b not in kernels:
            kernels[k[:-len(a)] + b] = kernels[k]
json.dump({"source": note or raw, "what": "per launch averages; traffic = dram__bytes_read.sum + dram__bytes_write.sum",
           "kernels": kernels}, open(out, "w"), indent=1)
print("wrote", out, len(kernels), "kernels")
