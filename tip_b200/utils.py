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


def _split_on_device(raw, raw_ptr_host, p):
    """train / test split of concatenated raw pairs on the GPU (csrc/edge_split.cu), drawing from the device-side
    numpy-compatible MT19937 stream of tip_b200.neg_sampling.  -> (train_idx, train_et, train_range, test_idx, test_et,
    test_range), bit for bit what src/utils.py:35-65 returns when numpy's global stream is in the same state."""
    import math
    from . import neg_sampling as ns
    from ._lib import TipbError, check, lib, ptr, stream
    from .ops import workspace
    if not 0.0 < p < 1.0:
        raise ValueError("process_edges on the device needs 0 < p < 1")
    dev, L = raw.device, lib()
    n_raw, n_rel = int(raw.shape[1]), len(raw_ptr_host) - 1
    q = 1.0 - p if p > 0.5 else p                 # numpy's legacy binomial runs the inversion sampler on min(p, 1-p)
    qn = math.exp(1 * math.log(1.0 - q))
    px2 = ((1 - 1 + 1) * q * qn) / (1 * (1.0 - q))
    with torch.cuda.device(dev):
        raw = raw.contiguous()
        raw_ptr = torch.tensor(raw_ptr_host, dtype=torch.long, device=dev)
        kept_scan = torch.empty(n_raw + 1, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = workspace(L.tipb_edge_split_workspace_bytes(n_raw), dev, "split")
        rng = ns._get_rng(dev)
        need = 2 * n_raw + 1024
        while True:
            rng.join()
            if not rng.valid or rng.n_new < need:
                rng.generate(max(need, rng.n_new))
                rng.valid = True
            check(L.tipb_edge_split_mask(ptr(rng.state), ptr(rng.words), 624 + rng.n_new, n_raw, qn, px2, int(p > 0.5),
                                         ptr(kept_scan), ptr(status), ptr(ws), ws.numel(), stream()), "edge_split_mask")
            code = int(status.item())
            if code == 0:
                rng.valid = False                  # the state moved on
                break
            if code & 2:
                raise TipbError("process_edges: a draw hit numpy's inversion-restart corner (U within 2^-53 of 1)")
            need *= 2
        k = int(kept_scan[n_raw].item())
        out = [torch.empty((2, 2 * k), dtype=torch.long, device=dev), torch.empty(2 * k, dtype=torch.long, device=dev),
               torch.empty((n_rel, 2), dtype=torch.long, device=dev),
               torch.empty((2, 2 * (n_raw - k)), dtype=torch.long, device=dev),
               torch.empty(2 * (n_raw - k), dtype=torch.long, device=dev), torch.empty((n_rel, 2), dtype=torch.long, device=dev)]
        check(L.tipb_edge_split_emit(ptr(raw), ptr(raw_ptr), n_raw, n_rel, ptr(kept_scan), k, *(ptr(t) for t in out),
                                     stream()), "edge_split_emit")
    return tuple(out)


def process_edges(raw_edge_list, p=0.9):
    """Bernoulli(p) train/test split per relation + bidirection + range lists (src/utils.py:35-65).
    CPU tensors: the global numpy stream exactly like the reference (one binomial call per relation).
    CUDA tensors: the same result from the device-side stream (tip_b200.neg_sampling.seed / set_state), computed on
    the GPU (csrc/edge_split.cu) -- the raw pairs never visit the host."""
    raw_edge_list = list(raw_edge_list)
    if raw_edge_list and all(torch.is_tensor(e) and e.is_cuda for e in raw_edge_list):
        ptr_host = [0]
        for e in raw_edge_list:
            ptr_host.append(ptr_host[-1] + int(e.shape[1]))
        return _split_on_device(torch.cat(raw_edge_list, dim=1), ptr_host, p)
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


def process_prot_edge(indices, p=0.9):
    """data/utils.py:212-229 for an int64 [2, E] tensor of protein-protein pairs (both directions present, as the scipy
    COO matrix lists them): keep one direction, Bernoulli(p) split, mirror.  -> (train_indices, test_indices).
    CUDA tensors are split on the device; CPU tensors draw from numpy's global stream like the reference."""
    indices = remove_bidirection(indices, None)
    if indices.is_cuda:
        res = _split_on_device(indices, [0, int(indices.shape[1])], p)
        return res[0], res[3]
    rd = np.random.binomial(1, p, indices.shape[1])
    return to_bidirection(indices[:, rd.nonzero()[0]]), to_bidirection(indices[:, (1 - rd).nonzero()[0]])


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
