"""BASELINE.json config 5: the full prediction tensor, all 645^2 drug pairs x 861 relations (358 M scores, 1.43 GB).
Times tipb_decoder_sweep with CUDA events (L2 flushed between launches) and spot-checks it against torch.
usage: python tools/ubench_sweep.py   (on a GPU box)"""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tip_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
n, r, dim = 645, 861, 16
z = torch.randn(n, dim, device=dev)
w = torch.randn(r, dim, device=dev) * 0.25
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.decoder_sweep(z, w, sigmoid=True)
    e1.record(); e1.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1) * 1e-3)
    if i < 12:
        del out
rel = torch.tensor([0, 17, 430, 860], device=dev)
ref = torch.sigmoid(torch.einsum("ik,rk,jk->rij", z, w[rel], z))
err = (out[rel] - ref).abs().max().item()
t = statistics.mean(ts)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
nbytes = out.numel() * 4
print(json.dumps({"workload": "decoder sweep 861 x 645 x 645 (BASELINE.json config 5)", "scores": out.numel(), "us": t * 1e6,
                  "scores_per_s": out.numel() / t, "write_GBps": nbytes / t / 1e9, "frac_of_measured_hbm": nbytes / t / 1e9 / peak,
                  "includes": "torch.empty of the 1.43 GB output inside the timed region", "max_abs_err_vs_torch": err}))
