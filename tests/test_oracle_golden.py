"""The oracle (oracle/*.py) against the golden vectors dumped from the reference's
own code (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import layout_oracle as lo
from oracle import neg_sampling_oracle as nso
from oracle import tip_oracle as to

RTOL, ATOL = 1e-4, 1e-6   # fp32 contract of BASELINE.json north_star


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale)


# ------------------------------------------------------------------ RNG / sampler (row N)
def test_mt19937_matches_numpy_stream():
    for seed in (1111, 0, 5, 2**32 - 1):
        rs = np.random.RandomState(seed)
        mt = nso.MT19937(seed)
        st = rs.get_state()
        assert np.array_equal(st[1], mt.key) and st[2] == mt.pos
        a = rs.randint(0, 2**32, size=2000, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(a, mt.words(2000))
        for n, k in ((645 * 645, 5000), (37 * 37, 100), (10**8, 3000), (1, 4), (2**20, 17)):
            assert np.array_equal(rs.choice(n, k), mt.choice(n, k))
        st = rs.get_state()
        assert np.array_equal(st[1], mt.key) and st[2] == mt.pos


def test_neg_sampling_matches_reference(golden_neg):
    for case in ("dense37", "drug645", "big10k", "tiny1"):
        g = {k.split("/", 1)[1]: v for k, v in golden_neg.items() if k.startswith(case + "/")}
        mt = nso.MT19937()
        mt.set_state(("MT19937", g["key0"], int(g["pos0"])))
        n = int(g["num_nodes"])
        out1 = nso.typed_negative_sampling(mt, g["pos"], n, g["range_list"])
        out2 = nso.typed_negative_sampling(mt, g["pos"], n, g["range_list"])
        assert np.array_equal(out1, g["neg_call1"]), case
        assert np.array_equal(out2, g["neg_call2"]), case
        assert np.array_equal(mt.key, g["key_end"]) and mt.pos == int(g["pos_end"])


def test_float32_row_semantics_differ_from_floordiv(golden_neg):
    # the contract is float32 true-division + trunc, not integer floor-div
    neg = golden_neg["big10k/neg_call1"]
    assert neg[0].max() <= 10000   # row == num_nodes can occur through rounding up
    perm_lo = neg[0] * 10000 + neg[1]
    assert (perm_lo // 10000 != neg[0]).sum() == 0  # self-consistent by construction
    # a value where the two semantics disagree exists in the key space
    p = np.int64(99999999)
    assert int(np.float32(p) / np.float32(10000)) != p // 10000


# ------------------------------------------------------------------ layout (row L)
def test_process_edges_matches_reference(golden_layout):
    g = golden_layout
    raw = [g[f"raw{i}"] for i in range(int(g["n_raw"]))]
    mt = nso.MT19937(1111)
    res = lo.process_edges(mt, raw, p=0.9)
    for got, name in zip(res, ["train_idx", "train_et", "train_range", "test_idx", "test_et", "test_range"]):
        assert np.array_equal(got, g[name]), name
    assert np.array_equal(lo.bernoulli_keep_mask(mt, 0.9, 64), g["binomial_head"])


def test_typed_csr_definition():
    rng = np.random.default_rng(0)
    n, r, e = 13, 4, 300
    ei = rng.integers(0, n, (2, e))
    et = rng.integers(0, r, e)
    for by in ("dst", "src"):
        c = lo.typed_csr(ei, et, n, r, by=by)
        node_row = 1 if by == "dst" else 0
        assert c["seg_ptr"][0] == 0 and c["seg_ptr"][-1] == e
        for s in range(len(c["seg_node"])):
            ids = c["eid"][c["seg_ptr"][s]:c["seg_ptr"][s + 1]]
            assert ids.size > 0
            assert (ei[node_row, ids] == c["seg_node"][s]).all() and (et[ids] == c["seg_rel"][s]).all()
            assert (np.diff(ids) > 0).all()   # stable: input order inside a segment
            assert np.array_equal(c["other"][c["seg_ptr"][s]:c["seg_ptr"][s + 1]], ei[1 - node_row, ids])
        assert np.array_equal(np.bincount(ei[node_row], minlength=n), c["deg"])
        assert c["node_ptr"][-1] == len(c["seg_node"])


# ------------------------------------------------------------------ operators (rows A2-A8)
def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.asarray(a)).to(dtype) if np.asarray(a).dtype.kind == "f" else torch.from_numpy(np.asarray(a))


def test_rgcn_forms_match_reference(golden_layers):
    g = golden_layers
    ei, et, rl = _t(g["data/dd_train_idx"]), _t(g["data/dd_train_et"]), _t(g["data/dd_train_range"])
    x = _t(g["rgcn2/x"]).requires_grad_(True)
    p = {k: _t(g[f"rgcn2/{k}"]).requires_grad_(True) for k in ("att", "basis", "root")}
    out = to.rgcn_conv_structural(x, ei, rl, p["att"], p["basis"], p["root"])
    close(out.detach(), g["rgcn2/out"])
    out.backward(_t(g["rgcn2/gout"]))
    close(x.grad, g["rgcn2/dx"])
    for k in p:
        close(p[k].grad, g[f"rgcn2/d_{k}"])
    with torch.no_grad():
        close(to.rgcn_conv_vectorized(x, ei, et, p["att"], p["basis"], p["root"]), g["rgcn2/out"])
        close(to.rgcn_conv_bmm(x, ei, et, p["att"], p["basis"], p["root"]), g["rgcn2/out"])
    # fp64 'truth' sits within the same tolerance of the fp32 reference
    out64 = to.rgcn_conv_vectorized(x.detach().double(), ei, et, *(p[k].detach().double() for k in ("att", "basis", "root")))
    close(out64, g["rgcn2/out"])
    # MyRGCNConv on shuffled edges
    perm = _t(g["rgcn1/perm"])
    q = {k: _t(g[f"rgcn1/{k}"]) for k in ("att", "basis", "root")}
    close(to.rgcn_conv_bmm(x.detach(), ei[:, perm], et[perm], q["att"], q["basis"], q["root"]), g["rgcn1/out"])


def test_rgcn_init_draw_order():
    import math
    torch.manual_seed(1111)
    p = to.init_rgcn_params(24, 12, 6, 5, after_relu=False)
    torch.manual_seed(1111)
    att = torch.empty(6, 5).normal_(std=1 / math.sqrt(5))
    assert torch.equal(att, p["att"])


def test_pp_hier_decoder_match_reference(golden_layers):
    g = golden_layers
    n_prot, n_drug = int(g["data/n_prot"]), int(g["data/n_drug"])
    ng = to.gcn_norm(_t(g["data/pp_train_indices"]), n_prot, torch.float32)
    xid = torch.eye(n_prot).to_sparse()
    out = to.pp_encoder(xid, ng, _t(g["pp/conv1.lin.weight"]), _t(g["pp/conv1.bias"]),
                        _t(g["pp/conv2.lin.weight"]), _t(g["pp/conv2.bias"]))
    close(out, g["pp/out"])
    oh = to.hier_conv(_t(g["hier/x"]), _t(g["data/dp_edge_index"]), _t(g["hier/weight"]), n_prot, n_drug)
    close(oh, g["hier/out"])
    ei, et = _t(g["data/dd_train_idx"]), _t(g["data/dd_train_et"])
    perm = _t(g["rgcn1/perm"])
    z, w = _t(g["dec/z"]), _t(g["dec/weight"])
    close(to.decoder(z, ei[:, perm], et[perm], w), g["dec/score"])
    close(to.decoder(z, ei[:, perm], et[perm], w, sigmoid=False), g["dec/value"])


def test_full_model_matches_reference(golden_layers):
    g = golden_layers
    n_prot, n_drug = int(g["data/n_prot"]), int(g["data/n_drug"])
    d = {k: _t(g[f"data/{k}"]) for k in ("dd_train_idx", "dd_train_et", "dd_train_range", "pp_train_indices",
                                         "dp_edge_index", "d_norm")}
    for mod in ("cat", "add"):
        pre = f"tip_{mod}/"
        for structural in (True, False):
            params = {n: _t(g[pre + "param/" + n]).requires_grad_(True) for n in to.TipOracle.param_names}
            orc = to.TipOracle(params, n_drug, n_prot, mod=mod, structural=structural)
            loss, z = orc.loss(d, _t(g[pre + "neg"]))
            close(z.detach(), g[pre + "z"])
            close(loss.detach(), g[pre + "loss"], rtol=1e-5)
            loss.backward()
            for n in to.TipOracle.param_names:
                close(params[n].grad, g[pre + "grad/" + n])
        # negatives: the oracle sampler reproduces what TIP.forward drew
        mt = nso.MT19937()
        mt.set_state(("MT19937", g[pre + "mt_key"], int(g[pre + "mt_pos"])))
        neg = nso.typed_negative_sampling(mt, g["data/dd_train_idx"], n_drug, g["data/dd_train_range"])
        assert np.array_equal(neg, g[pre + "neg"])


def test_ablation_operators_match_reference(golden_ablation):
    """the oracle's NNDecoder / HierEncoder / sparse-feature FMEncoder against the reference's own
    classes (tests/golden/ablation.npz, oracle/make_golden.py:golden_ablation)"""
    g = golden_ablation
    n_prot, n_drug = int(g["data/n_prot"]), int(g["data/n_drug"])
    ei, et = _t(g["data/dd_train_idx"]), _t(g["data/dd_train_et"])
    perm = _t(g["nn/perm"])
    w = {k: _t(g["nn/" + k]).requires_grad_(True) for k in ("w1_l1", "w1_l2", "w2_l1", "w2_l2")}
    z = _t(g["nn/z"]).requires_grad_(True)
    sc = to.nn_decoder(z, ei[:, perm], et[perm], w["w1_l1"], w["w1_l2"], w["w2_l1"], w["w2_l2"])
    close(sc.detach(), g["nn/score"])
    sc.backward(_t(g["nn/gscore"]))
    close(z.grad, g["nn/dz"])
    for k, v in w.items():
        close(v.grad, g["nn/d_" + k])
    # HierEncoder fed as test/pd_net.py:26 does
    feat = torch.cat([torch.eye(n_prot), torch.zeros(n_drug, n_prot)])
    embed, weight = _t(g["hier_enc/embed"]).requires_grad_(True), _t(g["hier_enc/hgcn.weight"]).requires_grad_(True)
    out = to.hier_encoder(feat, _t(g["data/dp_edge_index"]), _t(g["hier_enc/x_norm"]), embed, weight, n_prot, n_drug)
    close(out.detach(), g["hier_enc/out"])
    out.backward(_t(g["hier_enc/gout"]))
    close(embed.grad, g["hier_enc/d_embed"])
    close(weight.grad, g["hier_enc/d_hgcn.weight"])
    # FMEncoder with identity + mono side-effect drug features
    n_mono = int(g["fm_mono/n_mono"])
    idx = _t(g["fm_mono/feat_index"])
    d = {k: _t(g[f"data/{k}"]) for k in ("dd_train_idx", "dd_train_et", "dd_train_range", "pp_train_indices", "dp_edge_index")}
    d["d_norm"] = _t(g["fm_mono/d_norm"])
    d["d_feat"] = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n_drug, n_drug + n_mono))
    names = [n for n in to.TipOracle.param_names if n != "decoder.weight"]
    params = {n: _t(g["fm_mono/param/" + n[len("encoder."):]]).requires_grad_(True) for n in names}
    orc = to.TipOracle(params, n_drug, n_prot, mod="cat", structural=True)
    zz = orc.encode(d)
    close(zz.detach(), g["fm_mono/z"])
    zz.backward(_t(g["fm_mono/gz"]))
    for n in names:
        close(params[n].grad, g["fm_mono/grad/" + n[len("encoder."):]])


def test_eval_oracle_matches_reference_auprc_auroc_ap():
    """oracle/eval_oracle.py against tests/golden/eval.npz = outputs of the reference's own src/utils.py:86-93
    (scikit-learn) on seeded scores with heavy ties, saturated scores and a one-pair relation"""
    from oracle import eval_oracle as eo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval.npz"))
    rec = eo.record_by_relation(g["pos"], g["neg"], g["range_list"])
    np.testing.assert_allclose(rec, g["record"], rtol=0, atol=1e-14)
    # and against scikit-learn directly (it is what the reference calls), fresh random scores
    from tip_b200.utils import auprc_auroc_ap
    import torch
    rng = np.random.default_rng(5)
    s = rng.random(400).astype(np.float32).round(2)
    t = (rng.random(400) < 0.4).astype(np.float32)
    want = auprc_auroc_ap(torch.from_numpy(t), torch.from_numpy(s))
    np.testing.assert_allclose(eo.auprc_auroc_ap(t, s), want, rtol=0, atol=1e-14)
