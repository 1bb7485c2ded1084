"""Kernel-level timeline of ONE CUDA-graph replay of the TIP-cat step (CUPTI through torch.profiler):
start offset, duration and stream of every kernel -> which chain is the critical path.
usage: python tools/graph_trace.py [out.txt]   (on a GPU box)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers, neg_sampling as ns, optim
from torch.profiler import ProfilerActivity, profile

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
    return loss

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(4):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
model.embeddings = None
opt.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
_mp = int(os.environ.get("TIPB_BENCH_MAIN_PRIORITY", "0"))     # as bench.py: priority of the capture stream
with torch.cuda.graph(g, stream=torch.cuda.Stream(priority=_mp) if _mp != 0 else None):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
streams = {}
print("one graph replay: %d device activities, span %.1f us" % (len(evs), evs[-1].time_range.end - t0), file=out)
print("%9s %9s %8s  %-3s %s" % ("start_us", "end_us", "dur_us", "lane", "kernel"), file=out)
lanes = []   # greedy lane assignment = concurrent chains
for e in evs:
    s, en = e.time_range.start - t0, e.time_range.end - t0
    lane = None
    for i, busy_until in enumerate(lanes):
        if busy_until <= s + 0.5:
            lane = i
            break
    if lane is None:
        lanes.append(0.0)
        lane = len(lanes) - 1
    lanes[lane] = en
    print("%9.1f %9.1f %8.1f  %-3d %s" % (s, en, en - s, lane, e.name[:80]), file=out)
assert int(ns.last_status(dev)) == 0
