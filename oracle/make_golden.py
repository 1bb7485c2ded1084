"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):   python -m oracle.make_golden

What runs:
  * the reference's own `src/neg_sampling.py` and `src/utils.py`, unmodified;
  * the reference's own `src/layers.py`, unmodified, on top of `oracle/pyg_shim`
    (torch_geometric 2.0.1 is not installable here; the shim restates the slice
    of it that src/layers.py calls -- see oracle/pyg_shim/torch_geometric/__init__.py).
Nothing of the reference is copied into this repository: only inputs, parameter
values and outputs are stored.
"""
import os
import pickle
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TIP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "pyg_shim"))
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import src.layers as layers        # noqa: E402  (also seeds torch/numpy with 1111)
    import src.neg_sampling as ns      # noqa: E402
    import src.utils as utils          # noqa: E402
    layers.device = torch.device("cpu")  # src/layers.py:319 reads an undefined global `device`
    return layers, ns, utils


def _raw_relations(gen, n_drug, sizes):
    """per relation an int64 [2,k] array of (row<col) pairs in row-major (scipy COO) order."""
    iu = np.stack(np.triu_indices(n_drug, 1))
    rels = []
    for k in sizes:
        pick = np.sort(gen.choice(iu.shape[1], size=k, replace=False))
        rels.append(iu[:, pick].astype(np.int64))
    return rels


def golden_neg_sampling(ns):
    cases = {}
    gen = np.random.Generator(np.random.PCG64(7))
    # (name, num_nodes, per-relation undirected sizes, seed)
    specs = [("dense37", 37, [300, 0, 500, 11, 640], 1111),        # heavy collisions -> >=3 retry rounds
             ("drug645", 645, [2400, 755, 0, 9000, 252], 1111),   # polypharmacy node count
             ("big10k", 10000, [6000, 30000], 5),                  # float32 row rounding differs from floor-div
             ("tiny1", 3, [3], 99)]
    for name, n, sizes, seed in specs:
        rels = [np.concatenate([r, r[::-1]], axis=1) for r in _raw_relations(gen, n, sizes)]
        pos = np.concatenate(rels, axis=1)
        ends = np.cumsum([r.shape[1] for r in rels])
        rl = np.stack([ends - [r.shape[1] for r in rels], ends], axis=1).astype(np.int64)
        np.random.seed(seed)
        state0 = np.random.get_state()
        out1 = ns.typed_negative_sampling(torch.from_numpy(pos), n, torch.from_numpy(rl)).numpy()
        out2 = ns.typed_negative_sampling(torch.from_numpy(pos), n, torch.from_numpy(rl)).numpy()
        state2 = np.random.get_state()
        cases[name] = dict(num_nodes=n, pos=pos, range_list=rl, seed=seed, key0=state0[1], pos0=state0[2],
                           neg_call1=out1, neg_call2=out2, key_end=state2[1], pos_end=state2[2])
    flat = {f"{c}/{k}": np.asarray(v) for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "neg_sampling.npz"), **flat)
    return cases


def golden_layout(utils):
    gen = np.random.Generator(np.random.PCG64(11))
    raw = _raw_relations(gen, 50, [120, 0, 333, 7, 260])
    np.random.seed(1111)
    res = utils.process_edges([torch.from_numpy(r) for r in raw], p=0.9)
    names = ["train_idx", "train_et", "train_range", "test_idx", "test_et", "test_range"]
    flat = {n: v.numpy() for n, v in zip(names, res)}
    for i, r in enumerate(raw):
        flat[f"raw{i}"] = r
    flat["n_raw"] = np.array(len(raw))
    flat["binomial_head"] = np.random.binomial(1, 0.9, 64)  # stream position check after the call
    np.savez_compressed(os.path.join(OUT, "layout.npz"), **flat)


def _small_data_dict(utils, n_drug=40, n_prot=120, sizes=(90, 0, 150, 12, 260, 40), seed=3):
    gen = np.random.Generator(np.random.PCG64(seed))
    raw = _raw_relations(gen, n_drug - 1, list(sizes))      # drug n_drug-1 stays isolated (count clamp)
    np.random.seed(1111)
    d = {}
    (d["dd_train_idx"], d["dd_train_et"], d["dd_train_range"],
     d["dd_test_idx"], d["dd_test_et"], d["dd_test_range"]) = utils.process_edges([torch.from_numpy(r) for r in raw])
    pp = _raw_relations(gen, n_prot, [420])[0]
    d["pp_train_indices"] = torch.from_numpy(np.concatenate([pp, pp[::-1]], axis=1))
    n_dp = 70
    dp = np.stack([gen.integers(0, n_prot, n_dp), n_prot + gen.integers(0, n_drug // 2, n_dp)]).astype(np.int64)
    d["dp_edge_index"] = torch.from_numpy(dp)
    d["dp_range_list"] = torch.zeros((n_drug, 2))
    d["d_feat"], d["p_feat"] = utils.sparse_id(n_drug), utils.sparse_id(n_prot)
    d["n_drug"], d["n_prot"], d["n_dd_et"], d["n_drug_feat"] = n_drug, n_prot, len(sizes), n_drug
    d["d_norm"] = torch.ones(n_drug)
    return d


def golden_layers(layers, ns, utils):
    flat = {}
    d = _small_data_dict(utils)
    for k in ("dd_train_idx", "dd_train_et", "dd_train_range", "dd_test_idx", "dd_test_et", "dd_test_range",
              "pp_train_indices", "dp_edge_index", "d_norm"):
        flat[f"data/{k}"] = d[k].numpy()
    for k in ("n_drug", "n_prot", "n_dd_et"):
        flat[f"data/{k}"] = np.array(d[k])

    # ---- single operators (rows A2-A5, A7) with their own initialisers
    torch.manual_seed(1111)
    n_drug, n_prot, n_rel = d["n_drug"], d["n_prot"], d["n_dd_et"]
    ei, et, rl = d["dd_train_idx"], d["dd_train_et"], d["dd_train_range"]
    x = torch.randn(n_drug, 24, requires_grad=True)
    conv2 = layers.MyRGCNConv2(24, 12, n_rel, 5, after_relu=False)
    out = conv2(x, ei, et, rl)
    gout = torch.randn_like(out)
    out.backward(gout)
    flat.update({"rgcn2/x": x.detach().numpy(), "rgcn2/out": out.detach().numpy(), "rgcn2/gout": gout.numpy(),
                 "rgcn2/dx": x.grad.numpy()})
    for n, p in conv2.named_parameters():
        flat[f"rgcn2/{n}"], flat[f"rgcn2/d_{n}"] = p.detach().numpy(), p.grad.numpy()

    # MyRGCNConv (bmm form) on a shuffled edge order: no sort requirement
    perm = torch.randperm(ei.shape[1])
    conv1 = layers.MyRGCNConv(24, 12, n_rel, 5, after_relu=True)
    x1 = x.detach().clone().requires_grad_(True)
    out1 = conv1(x1, ei[:, perm], et[perm])
    out1.backward(gout)
    flat.update({"rgcn1/perm": perm.numpy(), "rgcn1/out": out1.detach().numpy(), "rgcn1/dx": x1.grad.numpy()})
    for n, p in conv1.named_parameters():
        flat[f"rgcn1/{n}"], flat[f"rgcn1/d_{n}"] = p.detach().numpy(), p.grad.numpy()

    pp = layers.PPEncoder(n_prot)
    xp = pp(d["p_feat"], d["pp_train_indices"])
    gp = torch.randn_like(xp)
    xp.backward(gp)
    flat.update({"pp/out": xp.detach().numpy(), "pp/gout": gp.numpy()})
    for n, p in pp.named_parameters():
        flat[f"pp/{n}"], flat[f"pp/d_{n}"] = p.detach().numpy(), p.grad.numpy()

    hc = layers.MyHierarchyConv(16, 10, n_prot, n_drug)
    xh = torch.cat([xp.detach(), torch.zeros(n_drug, 16)]).requires_grad_(True)
    oh = hc(xh, d["dp_edge_index"], d["dp_range_list"])
    gh = torch.randn_like(oh)
    oh.backward(gh)
    flat.update({"hier/x": xh.detach().numpy(), "hier/out": oh.detach().numpy(), "hier/gout": gh.numpy(),
                 "hier/dx": xh.grad.numpy(), "hier/weight": hc.weight.detach().numpy(),
                 "hier/d_weight": hc.weight.grad.numpy()})

    dec = layers.MultiInnerProductDecoder(12, n_rel)
    z = out.detach().clone().requires_grad_(True)
    sc = dec(z, ei[:, perm], et[perm])
    raw = dec(z, ei[:, perm], et[perm], sigmoid=False)
    gs = torch.randn_like(sc)
    sc.backward(gs)
    flat.update({"dec/z": z.detach().numpy(), "dec/score": sc.detach().numpy(), "dec/value": raw.detach().numpy(),
                 "dec/gscore": gs.numpy(), "dec/dz": z.grad.numpy(), "dec/weight": dec.weight.detach().numpy(),
                 "dec/d_weight": dec.weight.grad.numpy()})

    # ---- the whole model, through the reference's own TIP class (rows A6, A8)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "data_dict.pkl")
        with open(path, "wb") as f:
            pickle.dump(d, f)
        for mod, dims in (("cat", dict(prot_drug_dim=16, n_embed=48)), ("add", dict(prot_drug_dim=64, n_embed=64))):
            torch.manual_seed(1111)
            np.random.seed(1111)
            settings = layers.Setting(sp_rate=0.9, lr=0.01, n_hid1=32, n_hid2=16, num_base=32, **dims)
            model = layers.TIP(settings, torch.device("cpu"), mod=mod, data_path=path)
            state = np.random.get_state()
            neg = ns.typed_negative_sampling(model.data.dd_train_idx, n_drug, model.data.dd_train_range)
            np.random.set_state(state)
            model.train()
            loss = model()
            loss.backward()
            pre = f"tip_{mod}/"
            flat.update({pre + "loss": loss.detach().numpy(), pre + "z": model.embeddings.detach().numpy(),
                         pre + "neg": neg.numpy(), pre + "test_neg": model.test_neg_index.numpy(),
                         pre + "mt_key": state[1], pre + "mt_pos": np.array(state[2])})
            for n, p in model.named_parameters():
                flat[pre + "param/" + n], flat[pre + "grad/" + n] = p.detach().numpy().copy(), p.grad.numpy().copy()
            # three Adam steps (tip.py:21-30) to pin the training trajectory
            opt = torch.optim.Adam(model.parameters(), lr=settings.lr)
            np.random.set_state(state)
            losses = []
            for _ in range(3):
                opt.zero_grad()
                loss = model()
                losses.append(float(loss))
                loss.backward()
                opt.step()
            flat[pre + "loss_traj"] = np.array(losses)
    np.savez_compressed(os.path.join(OUT, "layers.npz"), **flat)


def golden_ablation(layers, utils):
    """NNDecoder, HierEncoder (src/layers.py:556-637) and an FMEncoder fed with general sparse
    drug features (identity + mono side-effect columns, data/utils.py:117-132), from the reference's own classes.
    Written to its own file so that the other fixtures (and their RNG history) stay byte-identical."""
    flat = {}
    d = _small_data_dict(utils)
    n_drug, n_prot, n_rel = d["n_drug"], d["n_prot"], d["n_dd_et"]
    ei, et = d["dd_train_idx"], d["dd_train_et"]
    for k in ("dd_train_idx", "dd_train_et", "dd_train_range", "pp_train_indices", "dp_edge_index", "d_norm"):
        flat[f"data/{k}"] = d[k].numpy()
    for k in ("n_drug", "n_prot", "n_dd_et"):
        flat[f"data/{k}"] = np.array(d[k])
    torch.manual_seed(2222)
    gen = np.random.Generator(np.random.PCG64(5))

    # ---- NNDecoder on a shuffled edge order
    perm = torch.randperm(ei.shape[1])
    dec = layers.NNDecoder(12, n_rel, l1_dim=16)
    z = torch.randn(n_drug, 12, requires_grad=True)
    sc = dec(z, ei[:, perm], et[perm])
    gs = torch.randn_like(sc)
    sc.backward(gs)
    flat.update({"nn/perm": perm.numpy(), "nn/z": z.detach().numpy(), "nn/score": sc.detach().numpy(), "nn/gscore": gs.numpy(),
                 "nn/dz": z.grad.numpy()})
    for n, p in dec.named_parameters():
        flat[f"nn/{n}"], flat[f"nn/d_{n}"] = p.detach().numpy(), p.grad.numpy()

    # ---- HierEncoder as test/pd_net.py:26,124 feeds it: dense identity rows for the sources, zero rows for the targets
    enc = layers.HierEncoder(n_prot, 32, 16, n_prot, n_drug)
    feat = torch.cat([utils.dense_id(n_prot), torch.zeros(n_drug, n_prot)], dim=0)
    x_norm = torch.from_numpy(gen.uniform(0.5, 2.0, n_prot + n_drug).astype(np.float32))
    out = enc(feat, d["dp_edge_index"], d["dp_range_list"], x_norm)
    go = torch.randn_like(out)
    out.backward(go)
    flat.update({"hier_enc/x_norm": x_norm.numpy(), "hier_enc/out": out.detach().numpy(), "hier_enc/gout": go.numpy()})
    for n, p in enc.named_parameters():
        flat[f"hier_enc/{n}"], flat[f"hier_enc/d_{n}"] = p.detach().numpy(), p.grad.numpy()

    # ---- FMEncoder with sparse drug features [n_drug, n_drug + n_mono] and a non-trivial d_norm
    n_mono = 30
    mono_r = gen.integers(0, n_drug, 150)
    mono_c = gen.integers(0, n_mono, 150)
    pairs = np.unique(np.stack([mono_r, mono_c]), axis=1)
    row = np.concatenate([np.arange(n_drug), pairs[0]])
    col = np.concatenate([np.arange(n_drug), pairs[1] + n_drug])
    d_feat = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col])), torch.ones(len(row)), (n_drug, n_drug + n_mono))
    d_norm = torch.sqrt(torch.sparse.sum(d_feat, dim=1).to_dense())
    fm = layers.FMEncoder(torch.device("cpu"), n_drug + n_mono, n_rel, n_prot, n_prot, n_drug, prot_drug_dim=16,
                          num_base=8, n_embed=48, n_hid1=32, n_hid2=16, mod="cat")
    zz = fm(d_feat, ei, et, d["dd_train_range"], d_norm, d["p_feat"], d["pp_train_indices"], d["dp_edge_index"],
            d["dp_range_list"])
    gz = torch.randn_like(zz)
    zz.backward(gz)
    flat.update({"fm_mono/feat_index": np.stack([row, col]).astype(np.int64), "fm_mono/n_mono": np.array(n_mono),
                 "fm_mono/d_norm": d_norm.numpy(), "fm_mono/z": zz.detach().numpy(), "fm_mono/gz": gz.numpy()})
    for n, p in fm.named_parameters():
        flat[f"fm_mono/param/{n}"], flat[f"fm_mono/grad/{n}"] = p.detach().numpy().copy(), p.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ablation.npz"), **flat)


def main():
    os.makedirs(OUT, exist_ok=True)
    layers, ns, utils = _import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "ablation":      # only the file added in round 2
        golden_ablation(layers, utils)
        return
    golden_neg_sampling(ns)
    golden_layout(utils)
    golden_layers(layers, ns, utils)
    golden_ablation(layers, utils)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
