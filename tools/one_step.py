"""Two eager TIP-cat steps on the polypharmacy shape (target for `ncu -k regex:...` captures).
usage: python tools/one_step.py [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tip_b200 import layers, neg_sampling as ns

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
data, _ = bench.make_data("polypharmacy")
torch.manual_seed(1111); ns.seed(1111, dev)
model = layers.TIP(bench.settings_for("cat"), dev, mod="cat", data=data)
opt = torch.optim.Adam(model.parameters(), lr=0.01, capturable=True, fused=True)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = model(check_status=False)
    loss.backward()
    opt.step()
    ns.join_prefetch(dev)
torch.cuda.synchronize()
print("loss", float(loss), "status", int(ns.last_status(dev)))
