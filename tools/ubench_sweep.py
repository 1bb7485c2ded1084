"""BASELINE.json config 5: the full prediction tensor, all 645^2 drug pairs x 861 relations (358 M scores, 1.43 GB).
Times tipb_decoder_sweep (C ABI, preallocated output: no allocation inside the timed region) with CUDA events, L2
flushed between launches, and checks it against a float64 evaluation on four relations.
TIPB_SWEEP_TC=0 selects the CUDA-core kernel (decoder.cu), default = the tcgen05 kernel (sweep_tc.cu).
usage: python tools/ubench_sweep.py   (on a GPU box)"""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tip_b200 import _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
n, r, dim = 645, 861, 16
SIG = int(os.environ.get("TIPB_SWEEP_SIG", "1"))
z = torch.randn(n, dim, device=dev)
w = torch.randn(r, dim, device=dev) * 0.25
out = torch.empty((r, n, n), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.tipb_decoder_sweep(z.data_ptr(), w.data_ptr(), n, r, dim, SIG, out.data_ptr(), _lib.stream()), "decoder_sweep")
    e1.record(); e1.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1) * 1e-3)
status = L.tipb_decoder_sweep_status()
# what a pure write stream of the same size reaches on this GPU (torch fill kernel): the practical ceiling of a
# write-only kernel, next to the copy figure (half reads, half writes) of MEASURED_PEAKS.json
tf = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out.fill_(0.5)
    e1.record(); e1.synchronize()
    if i >= 3:
        tf.append(e0.elapsed_time(e1) * 1e-3)
_lib.check(L.tipb_decoder_sweep(z.data_ptr(), w.data_ptr(), n, r, dim, SIG, out.data_ptr(), _lib.stream()), "decoder_sweep")
rel = torch.tensor([0, 17, 430, 860], device=dev)
ref = torch.einsum("ik,rk,jk->rij", z.double(), w[rel].double(), z.double())
ref = torch.sigmoid(ref) if SIG else ref
err = (out[rel].double() - ref).abs().max().item()
t = statistics.mean(ts)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
nbytes = out.numel() * 4
print(json.dumps({"workload": "decoder sweep 861 x 645 x 645 (BASELINE.json config 5)",
                  "kernel": "CUDA-core k_decoder_sweep_tiled" if os.environ.get("TIPB_SWEEP_TC") == "0" else "tcgen05 k_decoder_sweep_tc",
                  "scores": out.numel(), "us": t * 1e6, "us_min": min(ts) * 1e6,
                  "scores_per_s": out.numel() / t, "write_GBps": nbytes / t / 1e9, "frac_of_measured_hbm": nbytes / t / 1e9 / peak,
                  "timed": "C-ABI call on a preallocated output, CUDA events, L2 flushed between launches, mean of 10",
                  "fill_same_size_us": statistics.mean(tf) * 1e6, "fill_write_GBps": nbytes / statistics.mean(tf) / 1e9,
                  "frac_of_fill": statistics.mean(tf) / t,
                  "max_abs_err_vs_float64": err, "tc_status": status}))
