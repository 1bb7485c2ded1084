"""Two eager TIP-cat steps on the polypharmacy shape (target for `ncu -k regex:...` captures).
usage: python tools/one_step.py [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers, neg_sampling as ns, optim

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
torch.manual_seed(1111); ns.seed(1111, dev)
model = layers.TIP(bench.settings_for("cat"), dev, mod="cat", data=data)
opt = optim.Adam(model.parameters(), lr=0.01)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = model(check_status=False)
    loss.backward()
    opt.step()
    ns.join_prefetch(dev)
torch.cuda.synchronize()
print("loss", float(loss), "status", int(ns.last_status(dev)))
if os.environ.get("TIPB_DUMP_WORKLOAD"):
    import json
    from tip_b200 import ops
    d = model.data
    rl = d.dd_train_range
    p_dst = ops.cached_plan(d.dd_train_idx, d.n_drug, d.n_dd_et, by_src=False, edge_type=None, range_list=rl)
    p_src = ops.cached_plan(d.dd_train_idx, d.n_drug, d.n_dd_et, by_src=True, edge_type=None, range_list=rl)
    m = ns._membership(d.dd_train_idx, d.n_drug, rl)
    info = dict(E=int(d.dd_train_idx.shape[1]), n_drug=int(d.n_drug), n_prot=int(d.n_prot), n_rel=int(d.n_dd_et),
                S_dst=int(p_dst.field("counts")[0]), S_src=int(p_src.field("counts")[0]),
                S_neg=int(model._neg_plan.field("counts")[0]) if getattr(model, "_neg_plan", None) is not None else 0, E_pp=int(d.pp_train_indices.shape[1]),
                E_pd=int(d.dp_edge_index.shape[1]), mt_words=int(624 + ns._get_rng(dev).n_new), sum_l=int(m.sum_l),
                sum_w=int(m.sum_w))
    json.dump(info, open(os.environ["TIPB_DUMP_WORKLOAD"], "w"))
    print(info)
