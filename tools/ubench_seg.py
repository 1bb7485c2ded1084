"""Times tipb_seg_aggregate alone (CUDA events, L2 flushed between launches) on the polypharmacy D-D plan for
F = 64/32/16 and checks it against a torch index_add reference.  usage: python tools/ubench_seg.py"""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import _lib, ops

dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
idx = data["dd_train_idx"].to(dev); rl = data["dd_train_range"].to(dev).long()
n, r = int(data["n_drug"]), int(data["n_dd_et"])
plan = ops.cached_plan(idx, n, r, range_list=rl, by_src=False)
L = _lib.lib()
S = int(plan.field("counts")[0]); e = plan.n_entries
seg_ptr = plan.field("seg_ptr")[:S + 1].long(); other = plan.field("other")[:e].long()
seg_of = torch.repeat_interleave(torch.arange(S, device=dev), seg_ptr[1:] - seg_ptr[:-1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for f in (64, 32, 16):
    x = torch.randn(n, f, device=dev)
    out = torch.zeros(plan.seg_cap * f, dtype=torch.float32, device=dev)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.tipb_seg_aggregate(plan.buf.data_ptr(), e, n, r, x.data_ptr(), n, f, out.data_ptr(), _lib.stream()), "seg")
        e1.record(); e1.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ref = torch.zeros(S, f, device=dev, dtype=torch.float64).index_add_(0, seg_of, x.double()[other])
    err = (out[:S * f].view(S, f).double() - ref).abs().max().item()
    print("F=%3d  %.1f us (min %.1f)   max abs err vs fp64 %.2e   S=%d E=%d" % (f, statistics.mean(ts), min(ts), err, S, e))
