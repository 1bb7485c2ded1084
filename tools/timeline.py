"""Coarse per-phase timeline of one eager TIP-cat step (CUDA events on each stream).
usage: python tools/timeline.py   (on a GPU box)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers, neg_sampling as ns, optim, ops

dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
torch.manual_seed(1111); ns.seed(1111, dev)
model = layers.TIP(bench.settings_for("cat"), dev, mod="cat", data=data)
opt = optim.Adam(model.parameters(), lr=0.01)

def ev(stream=None):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream if stream is not None else torch.cuda.current_stream())
    return e

def step(marks=None):
    d = model.data
    opt.zero_grad(set_to_none=True)
    cur = torch.cuda.current_stream()
    if marks is not None: marks["t0"] = ev()
    model._side.wait_stream(cur)
    with torch.cuda.stream(model._side):
        if marks is not None: marks["s_begin"] = ev()
        neg = ns.typed_negative_sampling(d.dd_train_idx, d.n_drug, d.dd_train_range, check_status=False, out=model._neg_index)
        if marks is not None: marks["s_sampled"] = ev()
        model._neg_plan.build(neg, range_list=d.dd_train_range)
        if marks is not None: marks["s_plan"] = ev()
    model.embeddings = model._encode()
    if marks is not None: marks["m_encoded"] = ev()
    pos_plan = ops.cached_plan(d.dd_train_idx, d.n_drug, d.n_dd_et, range_list=d.dd_train_range, by_src=False, doubled=True, rel_major=True)
    loss = ops.bce_loss(model.embeddings, model.decoder.weight, pos_plan, model._neg_plan, neg_stream=model._side)
    if marks is not None: marks["m_loss"] = ev()
    loss.backward()
    if marks is not None: marks["m_bwd"] = ev()
    opt.step()
    if marks is not None: marks["m_adam"] = ev()
    ns.join_prefetch(dev)
    if marks is not None: marks["m_join"] = ev()
    return loss

for _ in range(4):
    model(); step()
torch.cuda.synchronize()
for rep in range(3):
    marks = {}
    step(marks)
    torch.cuda.synchronize()
    t0 = marks["t0"]
    print("rep", rep, {k: round(t0.elapsed_time(v), 3) for k, v in marks.items() if k != "t0"})
