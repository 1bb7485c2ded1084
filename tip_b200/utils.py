"""Host-side helpers with the names and results of the reference's src/utils.py.

These define the *edge layout contract* the kernels consume:
relation-sorted, bidirected edges plus a cumulative range_list.  They are one-off
CPU data preparation in the reference too (prepare.py) and draw from the same global
numpy stream, so results match `src/utils.py` bit for bit under the same seed.
"""
import numpy as np
import torch


def remove_bidirection(edge_index, edge_type):
    """keep one direction (row > col) of every bidirected pair  (src/utils.py:7-14)"""
    keep = (edge_index[0] > edge_index[1]).nonzero().view(-1)
    if edge_type is None:
        return edge_index[:, keep]
    return edge_index[:, keep], edge_type[keep]


def to_bidirection(edge_index, edge_type=None):
    """[pairs ..., mirrored pairs ...]  (src/utils.py:17-23)"""
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    if edge_type is None:
        return both
    return both, torch.cat([edge_type, edge_type])


def get_range_list(edge_list):
    """cumulative (start, end) per relation  (src/utils.py:26-32)"""
    sizes = torch.tensor([int(e.shape[1]) for e in edge_list], dtype=torch.long)
    ends = torch.cumsum(sizes, 0)
    return torch.stack([ends - sizes, ends], dim=1)


def process_edges(raw_edge_list, p=0.9):
    """Bernoulli(p) train/test split per relation + bidirection + range lists (src/utils.py:35-65).
    Uses the global numpy stream exactly like the reference (one binomial call per relation)."""
    split = {"train": ([], []), "test": ([], [])}
    for r, idx in enumerate(raw_edge_list):
        keep = np.random.binomial(1, p, idx.shape[1])
        for name, mask in (("train", keep), ("test", 1 - keep)):
            sel = mask.nonzero()[0]
            split[name][0].append(to_bidirection(idx[:, sel]))
            split[name][1].append(torch.full((2 * sel.size,), r, dtype=torch.long))
    out = []
    for name in ("train", "test"):
        edges, labels = split[name]
        out += [torch.cat(edges, dim=1), torch.cat(labels), get_range_list(edges)]
    return tuple(out)


def sparse_id(n):
    """n x n sparse COO identity, float32  (src/utils.py:68-75)"""
    i = torch.arange(n, dtype=torch.long)
    return torch.sparse_coo_tensor(torch.stack([i, i]), torch.ones(n, dtype=torch.float32), (n, n), check_invariants=False)


def dense_id(n):
    return torch.eye(n, dtype=torch.float32)


def is_sparse_identity(x):
    """True if x is a (coalescable) sparse COO identity matrix: features == node index."""
    if not x.is_sparse or x.shape[0] != x.shape[1]:
        return False
    x = x.coalesce()
    n = x.shape[0]
    idx, val = x.indices(), x.values()
    if idx.shape[1] != n:
        return False
    ar = torch.arange(n, device=idx.device)
    return bool((idx[0] == ar).all() and (idx[1] == ar).all() and (val == 1).all())


def auprc_auroc_ap(target_tensor, score_tensor):
    """evaluation metrics on the CPU through scikit-learn, as the reference does (src/utils.py:86-93)"""
    from sklearn import metrics
    y = target_tensor.detach().cpu().numpy()
    pred = score_tensor.detach().cpu().numpy()
    auroc, ap = metrics.roc_auc_score(y, pred), metrics.average_precision_score(y, pred)
    prec, rec, _ = metrics.precision_recall_curve(y, pred)
    return metrics.auc(rec, prec), auroc, ap
