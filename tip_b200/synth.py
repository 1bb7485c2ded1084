"""Synthetic tri-graph data of the polypharmacy shape (BASELINE.json configs).

There is no network access for the Decagon dataset on the benchmark box, so the benchmark and
the large tests run on generated graphs whose statistics follow the shipped data
(645 drugs, 19,081 proteins, relation sizes log-normal(7.73, 1.17), hub drugs/proteins), laid
out by the same `process_edges` contract as the reference's prepare.py.  The result is the
dict that prepare.py pickles into data/data_dict.pkl (keys of prepare.py:13-44).
"""
import numpy as np
import torch

from .utils import process_edges, sparse_id

POLYPHARMACY = dict(n_drug=645, n_prot=19081, n_rel=861, dd_undirected=4_600_000, pp_undirected=716_000,
                    pd_edges=18_600)
SCALED = dict(n_drug=10_000, n_prot=100_000, n_rel=4_000, dd_undirected=25_000_000, pp_undirected=4_000_000,
              pd_edges=300_000)


def _distinct_pairs(gen, n_nodes, count, prob):
    """`count` distinct (i<j) pairs with endpoints drawn from `prob`, in row-major (scipy COO) order."""
    count = int(min(count, n_nodes * (n_nodes - 1) // 2))
    keys = np.empty(0, dtype=np.int64)
    while keys.size < count:
        m = int((count - keys.size) * 1.6) + 16
        a, b = gen.choice(n_nodes, m, p=prob), gen.choice(n_nodes, m, p=prob)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        new = (lo * n_nodes + hi)[lo != hi]
        keys = np.unique(np.concatenate([keys, new]))
    if keys.size > count:
        keys = np.sort(gen.choice(keys, count, replace=False))
    return np.stack([keys // n_nodes, keys % n_nodes]).astype(np.int64)


def relation_sizes(gen, n_rel, total, lo=250, hi=30_000):
    ell = np.clip(gen.lognormal(7.73, 1.17, n_rel), lo, hi)
    sizes = np.maximum(np.rint(total * ell / ell.sum()).astype(np.int64), 1)
    return sizes


def make_tip_data(n_drug, n_prot, n_rel, dd_undirected, pp_undirected, pd_edges, seed=1111, sp_rate=0.9):
    gen = np.random.Generator(np.random.PCG64(seed))
    # ---- D-D: typed, undirected, hub drugs
    cap = n_drug * (n_drug - 1) // 2
    sizes = np.minimum(relation_sizes(gen, n_rel, dd_undirected, lo=min(250, max(dd_undirected // (4 * n_rel), 1))), cap)
    pop_d = gen.lognormal(0.0, 1.0, n_drug)
    pop_d /= pop_d.sum()
    raw = [torch.from_numpy(_distinct_pairs(gen, n_drug, k, pop_d)) for k in sizes]
    np.random.seed(1111)                           # the reference's split consumes the global numpy stream
    d = {}
    (d["dd_train_idx"], d["dd_train_et"], d["dd_train_range"],
     d["dd_test_idx"], d["dd_test_et"], d["dd_test_range"]) = process_edges(raw, p=sp_rate)
    # ---- P-P: undirected, heavy-tailed degrees, 0.9 kept for training, both directions, no self loops
    pop_p = gen.lognormal(0.0, 1.3, n_prot)
    pop_p /= pop_p.sum()
    pp = _distinct_pairs(gen, n_prot, pp_undirected, pop_p)
    pp = pp[:, gen.random(pp.shape[1]) < 0.9]
    d["pp_train_indices"] = torch.from_numpy(np.concatenate([pp, pp[::-1]], axis=1))
    # ---- P-D: 44% of the drugs have targets among 19% of the proteins, power-law counts, sorted by drug
    drugs = np.sort(gen.choice(n_drug, max(int(0.44 * n_drug), 1), replace=False))
    prots = gen.choice(n_prot, max(int(0.19 * n_prot), 1), replace=False)
    w = gen.pareto(1.2, drugs.size) + 1.0
    per_drug = np.maximum(np.rint(pd_edges * w / w.sum()).astype(np.int64), 1)
    per_drug = np.minimum(per_drug, prots.size)
    rows, cols = [], []
    for drug, k in zip(drugs, per_drug):
        cols.append(np.sort(gen.choice(prots, k, replace=False)))
        rows.append(np.full(k, drug))
    dp = np.stack([np.concatenate(cols), np.concatenate(rows) + n_prot]).astype(np.int64)
    d["dp_edge_index"] = torch.from_numpy(dp)
    counts = np.bincount(dp[1] - n_prot, minlength=n_drug)
    ends = np.cumsum(counts)
    d["dp_range_list"] = torch.tensor(np.stack([ends - counts, ends], axis=1), dtype=torch.float32)
    # ---- identity features (prepare.py:22-25)
    d["d_feat"], d["p_feat"] = sparse_id(n_drug), sparse_id(n_prot)
    d["n_drug"], d["n_prot"], d["n_dd_et"], d["n_drug_feat"] = n_drug, n_prot, n_rel, n_drug
    d["d_norm"] = torch.ones(n_drug)
    return d
