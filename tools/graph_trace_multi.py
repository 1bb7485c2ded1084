"""Kernel-level timeline of ONE CUDA-graph replay of the relation-sharded TIP-cat step on rank 0 of an N-rank run
(CUPTI through torch.profiler): which chain is the critical path once the relations are sharded.
usage: torchrun --nproc-per-node N tools/graph_trace_multi.py out.txt [--shape scaled] [--model dd]"""
import datetime, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import neg_sampling as ns, optim, parallel
from torch.profiler import ProfilerActivity, profile

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
shape = "scaled" if "--shape" in sys.argv and sys.argv[sys.argv.index("--shape") + 1] == "scaled" else "polypharmacy"
model_kind = "dd" if "--model" in sys.argv and sys.argv[sys.argv.index("--model") + 1] == "dd" else "tip"
data, _ = bench.make_data(shape)
torch.manual_seed(1111); ns.seed(1111, dev)
coll = parallel._Collective(world, sampler_group=dist.new_group() if world > 1 else None)
cls = parallel.ShardedDDNet if model_kind == "dd" else parallel.ShardedTIP
model = cls(bench.settings_for("cat"), dev, mod="cat", data=data, rank=rank, world=world, collective=coll,
            defer_loss_reduce=world > 1)
opt = optim.Adam(model.parameters(), lr=0.01)
one = torch.ones((), device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    loss = model(check_status=False)
    loss.backward(one)
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
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # keep the LAST replay: events after the last k_adam_tick but one
    ticks = [i for i, e in enumerate(evs) if "k_adam_tick" in e.name]
    if len(ticks) >= 2:
        evs = evs[ticks[-2] + 1:]
    t0 = evs[0].time_range.start
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else sys.stdout
    print("rank 0 of %d, one graph replay (steady state): %d device activities, span %.1f us" % (world, len(evs), evs[-1].time_range.end - t0), file=out)
    print("%9s %9s %8s  %-3s %s" % ("start_us", "end_us", "dur_us", "lane", "kernel"), file=out)
    lanes = []
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
    out.flush()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
