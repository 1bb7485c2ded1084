"""R-GCN layer micro-benchmark on the polypharmacy D-D graph: per-kernel device time (CUPTI) of one forward + backward of
MyRGCNConv2(64 -> 32) and (32 -> 16), 861 relations, 32 bases.  TIPB_RGCN_TC / TIPB_RGCN_TC_DBG select kernel variants.
usage: python tools/ubench_rgcn.py   (on a GPU box)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
ei, et, rl = (data[k].to(dev) for k in ("dd_train_idx", "dd_train_et", "dd_train_range"))
torch.manual_seed(0)
res = {}
for fi, fo in ((64, 32), (32, 16)):
    conv = layers.MyRGCNConv2(fi, fo, int(data["n_dd_et"]), 32, after_relu=False).to(dev)
    x = torch.randn(int(data["n_drug"]), fi, device=dev, requires_grad=True)
    g = torch.randn(int(data["n_drug"]), fo, device=dev)
    for _ in range(3):
        conv(x, ei, et, rl).backward(g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            conv(x, ei, et, rl).backward(g)
        torch.cuda.synchronize()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "tipb::k_" in e.name:
            name = e.name.split("tipb::")[1].split("(")[0]
            k = res.setdefault("%d->%d %s" % (fi, fo, name), [0.0, 0])
            k[0] += e.time_range.end - e.time_range.start
            k[1] += 1
print(json.dumps({"TIPB_RGCN_TC": os.environ.get("TIPB_RGCN_TC"), "TIPB_RGCN_TC_DBG": os.environ.get("TIPB_RGCN_TC_DBG"),
                  "us": {k: round(v[0] / v[1], 1) for k, v in res.items() if "node" in k or "aggregate" in k or "k_rd" in k or "k_basis" in k or "rel_reduce" in k or "atb" in k}}))
