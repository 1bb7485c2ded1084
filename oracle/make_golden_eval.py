"""TEST INFRASTRUCTURE ONLY.  Runs the reference's own evaluation helper (src/utils.py:86-93, which calls scikit-learn)
on seeded scores and writes tests/golden/eval.npz.   usage: python oracle/make_golden_eval.py   (needs /root/reference)"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from src.utils import auprc_auroc_ap  # noqa: E402  (the reference's function, unmodified)

rng = np.random.default_rng(2024)
sizes = np.array([1, 2, 7, 300, 1500, 64, 5000, 33, 257, 256, 255, 1024], dtype=np.int64)
ends = np.cumsum(sizes)
rl = np.stack([ends - sizes, ends], axis=1)
e = int(ends[-1])
pos = (1 / (1 + np.exp(-rng.normal(0.8, 1.5, e)))).astype(np.float32)
neg = (1 / (1 + np.exp(-rng.normal(-0.5, 1.5, e)))).astype(np.float32)
# heavy ties: quantised scores in some relations, saturated scores in another, identical pos/neg scores in one
for r in (3, 6):
    a, b = rl[r]
    pos[a:b] = np.round(pos[a:b] * 20) / 20
    neg[a:b] = np.round(neg[a:b] * 20) / 20
a, b = rl[4]
pos[a:a + 700] = 1.0
neg[a:a + 90] = 1.0
neg[b - 200:b] = 0.0
a, b = rl[7]
pos[a:b] = 0.5
neg[a:b] = 0.5
rec = np.zeros((3, len(sizes)))
for r, (a, b) in enumerate(rl):
    score = torch.cat([torch.from_numpy(pos[a:b]), torch.from_numpy(neg[a:b])])
    target = torch.cat([torch.ones(b - a), torch.zeros(b - a)])
    rec[:, r] = auprc_auroc_ap(target, score)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eval.npz")
np.savez_compressed(out, pos=pos, neg=neg, range_list=rl, record=rec)
print(rec.T)
