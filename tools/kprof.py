"""Per-kernel time table of eager TIP-cat steps (CUPTI through torch.profiler; not under ncu).
usage: python tools/kprof.py [steps]   (on a GPU box)"""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers, neg_sampling as ns, optim
from torch.profiler import ProfilerActivity, profile

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
torch.manual_seed(1111); ns.seed(1111, dev)
model = layers.TIP(bench.settings_for("cat"), dev, mod="cat", data=data)
opt = optim.Adam(model.parameters(), lr=0.01)

def step():
    opt.zero_grad(set_to_none=True)
    loss = model(check_status=False)
    loss.backward()
    opt.step()
    ns.join_prefetch(dev)

for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name[:70]
        t, n = agg.get(name, (0.0, 0))
        agg[name] = (t + ev.device_time, n + 1)
tot = sum(t for t, _ in agg.values()) / steps
print("sum of kernel time per step: %.1f us" % tot)
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%9.1f us  x%3d  %s" % (t / steps, n // steps, name))
assert int(ns.last_status(dev)) == 0
