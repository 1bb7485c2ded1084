"""One-off: ONE full TIP-cat training step of the reference's CPU path (structural oracle = op-for-op restatement of
src/layers.py + src/neg_sampling.py, incl. the 861-iteration loops and autograd's O(R*E*F) slice backward) on the FULL
benchmark workload (all 861 relations, 8.28 M directed edges) -- the same-config anchor for the sub-sampled
`bench.py --impl reference` arm.  Takes minutes; not part of the default bench run.
usage: python tools/cpu_full_step.py [out.json] [mod]"""
import json
import os
import platform
import resource
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "cpu_full_step.json")
mod = sys.argv[2] if len(sys.argv) > 2 else "cat"
torch.set_num_threads(os.cpu_count() or 1)
data, workload = bench.make_data("polypharmacy", mod)
n_rel = int(data["n_dd_et"])
step, e_s, n_rel_used = bench.cpu_reference_step_factory(data, mod, n_rel)
assert n_rel_used == n_rel and e_s == int(data["dd_train_idx"].shape[1])
phases = {}
t0 = time.perf_counter()
loss = step(phases)
dt = time.perf_counter() - t0
cpu_model = ""
try:
    for line in open("/proc/cpuinfo"):
        if line.startswith("model name"):
            cpu_model = line.split(":", 1)[1].strip()
            break
except OSError:
    pass
rec = {"what": "ONE full TIP-%s training step (neg sampling + fwd + bwd + Adam) of the structural oracle on the full "
               "bench workload; first step (includes the one-time GCN normalisation)" % mod,
       "workload": workload, "relations": n_rel, "directed_dd_edges": e_s, "seconds": dt, "phases_s": phases,
       "typed_edge_msgs_per_s": 4.0 * e_s / dt, "cores": torch.get_num_threads(), "cpu_model": cpu_model,
       "machine": platform.node(), "peak_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0,
       "loss": loss, "torch": torch.__version__}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec))
