"""Property tests of the oracle itself (CPU only; SURVEY.md section 4 test plan): the three R-GCN forms agree on
hypothesis-generated graphs with the edge cases the domain has (empty relation, isolated node -> count clamp to 1,
duplicate edges, self pairs), autograd of the oracle passes `gradcheck` in float64, the numpy MT19937 / typed
negative sampler equal numpy and the reference's own function on generated inputs, and the evaluation oracle equals
scikit-learn on generated score vectors with ties.  A wrong oracle would make every GPU parity claim void."""
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import eval_oracle as eo
from oracle import layout_oracle as lo
from oracle import neg_sampling_oracle as nso
from oracle import tip_oracle as to

REF = "/root/reference"
SET = settings(max_examples=25, deadline=None, derandomize=True)


@st.composite
def typed_graphs(draw, max_nodes=9, max_rel=4, max_edges=30):
    n = draw(st.integers(2, max_nodes))
    r = draw(st.integers(1, max_rel))
    sizes = [draw(st.integers(0, max_edges // r)) for _ in range(r)]      # empty relations allowed
    e = sum(sizes)
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, max(n - 2, 0)), min_size=e, max_size=e))   # node n-1 never a target: isolated
    ei = torch.tensor([src, dst], dtype=torch.long).reshape(2, e)
    et = torch.repeat_interleave(torch.arange(r), torch.tensor(sizes))
    ends = np.cumsum(sizes)
    rl = torch.tensor(np.stack([ends - sizes, ends], axis=1), dtype=torch.long).reshape(r, 2)
    return n, r, ei, et, rl


@SET
@given(typed_graphs(), st.integers(0, 2 ** 16))
def test_rgcn_forms_agree_on_generated_graphs(g, seed):
    n, r, ei, et, rl = g
    gen = torch.Generator().manual_seed(seed)
    fi, fo, nb = 3, 2, 2
    x = torch.randn(n, fi, generator=gen, dtype=torch.float64)
    att = torch.randn(r, nb, generator=gen, dtype=torch.float64)
    basis = torch.randn(nb, fi, fo, generator=gen, dtype=torch.float64)
    root = torch.randn(fi, fo, generator=gen, dtype=torch.float64)
    a = to.rgcn_conv_structural(x, ei, rl, att, basis, root)
    b = to.rgcn_conv_bmm(x, ei, et, att, basis, root)
    c = to.rgcn_conv_vectorized(x, ei, et, att, basis, root)
    torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(a, c, rtol=1e-12, atol=1e-12)
    # a node without incoming edges only keeps its root term (mean over zero edges = 0, count clamped to 1)
    torch.testing.assert_close(a[n - 1], x[n - 1] @ root, rtol=1e-12, atol=1e-12)
    # MyRGCNConv accepts edge types in any order: permuting the edges does not change the result
    if ei.shape[1] > 1:
        perm = torch.randperm(ei.shape[1], generator=gen)
        torch.testing.assert_close(to.rgcn_conv_bmm(x, ei[:, perm], et[perm], att, basis, root), a, rtol=1e-10, atol=1e-10)


def _tiny_graph():
    ei = torch.tensor([[0, 1, 2, 2, 3, 0, 1, 4], [1, 0, 1, 3, 2, 2, 1, 0]])
    et = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2])
    rl = torch.tensor([[0, 3], [3, 5], [5, 8]])
    return 5, ei, et, rl


def test_oracle_autograd_passes_gradcheck():
    n, ei, et, rl = _tiny_graph()
    gen = torch.Generator().manual_seed(3)
    mk = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64, requires_grad=True)
    x, att, basis, root = mk(n, 3), mk(3, 2), mk(2, 3, 2), mk(3, 2)
    assert torch.autograd.gradcheck(lambda *a: to.rgcn_conv_structural(a[0], ei, rl, *a[1:]), (x, att, basis, root))
    assert torch.autograd.gradcheck(lambda *a: to.rgcn_conv_vectorized(a[0], ei, et, *a[1:]), (x, att, basis, root))
    # hierarchy conv: 3 sources, 2 targets
    xh, wh = mk(5, 2), mk(2, 3)
    eh = torch.tensor([[0, 1, 2, 2], [3, 3, 4, 3]])
    assert torch.autograd.gradcheck(lambda a, b: to.hier_conv(a, eh, b, 3, 2), (xh, wh))
    # GCN layer over the normalised graph (self loops added once, duplicates kept)
    pp = torch.tensor([[0, 1, 1, 2, 3, 3], [1, 0, 2, 1, 3, 0]])
    norm = to.gcn_norm(pp, 4, torch.float64)
    xg, wg, bg = mk(4, 3), mk(2, 3), mk(2)
    assert torch.autograd.gradcheck(lambda a, b, c: to.gcn_conv(a, norm, b, c), (xg, wg, bg))
    # decoder + loss
    z, w = mk(n, 4), mk(3, 4)
    neg = torch.tensor([[4, 3, 0, 1, 2, 3, 4, 0], [4, 0, 2, 2, 1, 0, 3, 1]])
    assert torch.autograd.gradcheck(lambda a, b: to.tip_loss(to.decoder(a, ei, et, b), to.decoder(a, neg, et, b)), (z, w))


def test_decoder_is_symmetric_and_loss_matches_the_reference_formula():
    n, ei, et, _ = _tiny_graph()
    gen = torch.Generator().manual_seed(9)
    z, w = torch.randn(n, 4, generator=gen), torch.randn(3, 4, generator=gen)
    s = to.decoder(z, ei, et, w)
    torch.testing.assert_close(s, to.decoder(z, ei.flip(0), et, w))          # DistMult: z_i^T diag(w_r) z_j
    neg = to.decoder(z, ei.flip(1), et, w)
    want = -torch.log(s + 1e-13).mean() - torch.log(1 - neg + 1e-13).mean()   # src/layers.py:338-340
    torch.testing.assert_close(to.tip_loss(s, neg), want)


@SET
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 700))
def test_mt19937_oracle_equals_numpy_for_any_seed(seed, count):
    mt = nso.MT19937(seed)
    rs = np.random.RandomState(seed)
    assert np.array_equal(mt.choice(645 * 645, count), rs.choice(645 * 645, count))
    st_np = rs.get_state()
    assert np.array_equal(mt.key, st_np[1]) and mt.pos == st_np[2]


def _reference_neg_sampling():
    if not os.path.isdir(os.path.join(REF, "src")):
        pytest.skip("the reference tree is only present in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.neg_sampling import typed_negative_sampling      # the reference's own function, unmodified
    return typed_negative_sampling


@settings(max_examples=15, deadline=None, derandomize=True)
@given(st.integers(3, 40), st.lists(st.integers(0, 60), min_size=1, max_size=5), st.integers(0, 2 ** 31 - 1))
def test_neg_sampling_oracle_equals_the_live_reference(n, sizes, seed):
    """dense relations force the retry loop (and its indexing quirk, src/neg_sampling.py:12-16) to run"""
    ref_fn = _reference_neg_sampling()
    rng = np.random.default_rng(seed)
    sizes = [min(s, n * n - 1) for s in sizes]
    e = sum(sizes)
    pos = rng.integers(0, n, (2, e)).astype(np.int64)
    ends = np.cumsum(sizes)
    rl = np.stack([ends - sizes, ends], axis=1).astype(np.int64)
    np.random.seed(seed % (2 ** 32))
    want = ref_fn(torch.from_numpy(pos), n, torch.from_numpy(rl)).numpy()
    mt = nso.MT19937(seed % (2 ** 32))
    got = nso.typed_negative_sampling(mt, pos, n, rl)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    st_np = np.random.get_state()
    assert np.array_equal(mt.key, st_np[1]) and mt.pos == st_np[2]


@SET
@given(typed_graphs(max_nodes=12, max_rel=5, max_edges=60), st.booleans(), st.booleans())
def test_typed_csr_definition_properties(g, by_src, rel_major):
    """the index structure the CUDA builder must reproduce bit for bit: a stable grouping of the edge ids"""
    n, r, ei, et, rl = g
    csr = lo.typed_csr(ei.numpy(), et.numpy(), n, r, by="src" if by_src else "dst", rel_major=rel_major)
    e = ei.shape[1]
    eid, other, seg_ptr = csr["eid"], csr["other"], csr["seg_ptr"]
    assert sorted(eid.tolist()) == list(range(e))                        # a permutation of the edges
    node_col, other_col = (0, 1) if by_src else (1, 0)
    assert np.array_equal(other, ei.numpy()[other_col][eid])
    assert seg_ptr[0] == 0 and seg_ptr[-1] == e and np.all(np.diff(seg_ptr) > 0)      # no empty segment
    keys = []
    for s in range(len(seg_ptr) - 1):
        ids = eid[seg_ptr[s]:seg_ptr[s + 1]]
        assert np.all(np.diff(ids) > 0)                                  # input order inside a segment (stable)
        assert np.all(ei.numpy()[node_col][ids] == csr["seg_node"][s]) and np.all(et.numpy()[ids] == csr["seg_rel"][s])
        keys.append((csr["seg_rel"][s], csr["seg_node"][s]) if rel_major else (csr["seg_node"][s], csr["seg_rel"][s]))
    assert keys == sorted(set(keys))                                      # segments strictly ordered, no duplicates


@SET
@given(st.lists(st.tuples(st.integers(0, 20), st.booleans()), min_size=2, max_size=120))
def test_eval_oracle_equals_scikit_learn_with_ties(items):
    from sklearn import metrics
    score = np.array([s for s, _ in items], dtype=np.float32) / 20.0      # few distinct values: many ties
    y = np.array([t for _, t in items], dtype=np.float64)
    if y.min() == y.max():
        assert all(np.isnan(v) for v in eo.auprc_auroc_ap(y, score))
        return
    auprc, auroc, ap = eo.auprc_auroc_ap(y, score)
    p, r, _ = metrics.precision_recall_curve(y, score)
    np.testing.assert_allclose(auprc, metrics.auc(r, p), rtol=0, atol=1e-12)
    np.testing.assert_allclose(auroc, metrics.roc_auc_score(y, score), rtol=0, atol=1e-12)
    np.testing.assert_allclose(ap, metrics.average_precision_score(y, score), rtol=0, atol=1e-12)
    # AUROC of the complemented labels mirrors around 1/2; a monotone map of the scores changes nothing
    np.testing.assert_allclose(eo.auprc_auroc_ap(1 - y, score)[1], 1 - auroc, rtol=0, atol=1e-12)
    np.testing.assert_allclose(eo.auprc_auroc_ap(y, np.exp(score))[1], auroc, rtol=0, atol=1e-12)


def test_parameter_names_and_order_match_the_reference_state_dict():
    """parameter order drives state_dict / Adam order; checkpoints of the reference must carry over.  The golden file
    holds TIP.named_parameters() of the reference's own class as executed (oracle/make_golden.py)."""
    from tip_b200 import layers
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "layers.npz"))
    golden = tuple(k[len("tip_cat/param/"):] for k in g.files if k.startswith("tip_cat/param/"))
    torch.manual_seed(0)
    enc = layers.FMEncoder("cpu", 30, 4, 50, 50, 30, 16, 32, 48, 32, 16, mod="cat")
    names = ["encoder." + n for n, _ in enc.named_parameters()] + ["decoder.weight"]
    assert tuple(names) == golden == to.TipOracle.param_names
    shapes = {n: tuple(p.shape) for n, p in enc.named_parameters()}
    assert shapes["rgcn1.basis"] == (32, 64, 32) and shapes["rgcn1.att"] == (4, 32) and shapes["rgcn1.root"] == (64, 32)
    assert shapes["rgcn2.basis"] == (32, 32, 16) and shapes["embed"] == (30, 48) and shapes["hgcn.weight"] == (16, 16)
    assert shapes["pp_encoder.conv1.lin.weight"] == (32, 50) and shapes["pp_encoder.conv2.lin.weight"] == (16, 32)


def test_fullscale_oracle_form_equals_the_structural_form():
    """The benched workload (861 relations, 8.28 M edges) is checked on the GPU against TipOracle(sparse=True) in fp64
    (tests/test_gpu_parity.py::test_benched_workload_parity), because the reference's own op order costs O(R*E*F) in
    backward.  Here that form is cross-checked against the STRUCTURAL form (the reference's op order, per-relation loops)
    on a 48-relation instance with the model's real widths (64 -> 32 -> 16, 32 bases): loss, z and all 13 gradients."""
    from oracle import neg_sampling_oracle as nso
    from oracle import tip_oracle as to
    from tip_b200 import synth
    data = synth.make_tip_data(n_drug=300, n_prot=900, n_rel=48, dd_undirected=60_000, pp_undirected=5_000, pd_edges=700,
                               seed=21)
    n_drug, n_prot, n_rel = data["n_drug"], data["n_prot"], data["n_dd_et"]
    gen = torch.Generator().manual_seed(5)
    for mod, pd, ne in (("cat", 16, 48), ("add", 64, 64)):
        f0 = pd + ne if mod == "cat" else ne
        p = {"encoder.embed": torch.randn(n_drug, ne, generator=gen),
             "encoder.pp_encoder.conv1.bias": torch.randn(32, generator=gen) * 0.1,
             "encoder.pp_encoder.conv1.lin.weight": torch.randn(32, n_prot, generator=gen) * 0.05,
             "encoder.pp_encoder.conv2.bias": torch.randn(16, generator=gen) * 0.1,
             "encoder.pp_encoder.conv2.lin.weight": torch.randn(16, 32, generator=gen) * 0.3,
             "encoder.hgcn.weight": torch.randn(16, pd, generator=gen) * 0.25,
             "decoder.weight": torch.randn(n_rel, 16, generator=gen) * 0.25}
        for name, (fi, fo, relu) in (("rgcn1", (f0, 32, False)), ("rgcn2", (32, 16, True))):
            for k, v in to.init_rgcn_params(fi, fo, n_rel, 32, relu, gen).items():
                p[f"encoder.{name}.{k}"] = v
        cpu = {k: v for k, v in data.items() if torch.is_tensor(v) and not v.is_sparse}
        neg = torch.from_numpy(nso.typed_negative_sampling(nso.MT19937(3), cpu["dd_train_idx"].numpy(), n_drug,
                                                           cpu["dd_train_range"].numpy()))
        results = []
        for kw in (dict(structural=True), dict(structural=False, sparse=True)):
            q = to.to_dtype(p, torch.float64, requires_grad=True)
            loss, z = to.TipOracle(q, n_drug, n_prot, mod=mod, **kw).loss(cpu, neg)
            loss.backward()
            results.append((loss.detach(), z.detach(), {k: v.grad for k, v in q.items()}))
        (l0, z0, g0), (l1, z1, g1) = results
        assert len(g0) == 13
        torch.testing.assert_close(l1, l0, rtol=1e-12, atol=0)
        torch.testing.assert_close(z1, z0, rtol=1e-10, atol=1e-13)
        for k in g0:
            torch.testing.assert_close(g1[k], g0[k], rtol=1e-9, atol=1e-12 * float(g0[k].abs().max()), msg=lambda m: k + ": " + m)


def test_only_proteins_with_a_drug_edge_need_the_second_gcn_layer():
    """tip_b200.layers.FMEncoder runs PPEncoder.conv2 only over the P-P edges INTO proteins that are the source of a
    P->D edge, with the normalisation of the whole graph.  In the reference's formulation (oracle restatement of
    src/layers.py:391-395 and :229-242, float64): the hierarchy conv's output and the gradients of every parameter are
    the same whether conv2 aggregates over all edges or over that subset -- the other rows are never read."""
    from oracle import tip_oracle as to
    rng = np.random.default_rng(4)
    n_prot, n_drug = 300, 40
    pp = rng.integers(0, n_prot, size=(2, 1500))
    pp = np.concatenate([pp, pp[::-1]], axis=1)
    src = rng.choice(n_prot, size=60, replace=False)[rng.integers(0, 60, size=150)]
    dp = np.stack([src, n_prot + rng.integers(0, n_drug, size=150)])
    pp_t, dp_t = torch.from_numpy(pp), torch.from_numpy(dp)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(n_prot, 24, generator=gen, dtype=torch.float64)

    def run(restrict):
        w1 = (torch.randn(32, 24, generator=torch.Generator().manual_seed(1), dtype=torch.float64) * 0.2).requires_grad_(True)
        b1 = (torch.randn(32, generator=torch.Generator().manual_seed(2), dtype=torch.float64) * 0.1).requires_grad_(True)
        w2 = (torch.randn(16, 32, generator=torch.Generator().manual_seed(3), dtype=torch.float64) * 0.2).requires_grad_(True)
        b2 = (torch.randn(16, generator=torch.Generator().manual_seed(4), dtype=torch.float64) * 0.1).requires_grad_(True)
        wh = (torch.randn(16, 8, generator=torch.Generator().manual_seed(5), dtype=torch.float64) * 0.3).requires_grad_(True)
        row, col, w = to.gcn_norm(pp_t, n_prot, torch.float64)
        h = torch.relu(to.gcn_conv(x, (row, col, w), w1, b1))
        if restrict:
            read = torch.zeros(n_prot, dtype=torch.bool)
            read[dp_t[0]] = True
            keep = read[col]                      # edges (incl. the self loops gcn_norm added) INTO a protein that is read
            assert 0 < int(keep.sum()) < keep.numel()
            row, col, w = row[keep], col[keep], w[keep]
        x2 = to.gcn_conv(h, (row, col, w), w2, b2)
        out = to.hier_conv(torch.cat([x2, torch.zeros(n_drug, 16, dtype=torch.float64)]), dp_t, wh, n_prot, n_drug)
        (out * torch.arange(1, out.numel() + 1, dtype=torch.float64).reshape(out.shape)).sum().backward()
        return out.detach(), [p.grad for p in (w1, b1, w2, b2, wh)]

    out_full, g_full = run(False)
    out_sub, g_sub = run(True)
    torch.testing.assert_close(out_sub, out_full, rtol=0, atol=0)
    for a, b in zip(g_sub, g_full):
        torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-14)


def test_gcn_and_hierarchy_oracle_against_dense_textbook_forms():
    """Independent witness for the third-party half of rows A4 / A5 (PyG's GCNConv is not installable here): the
    oracle's gcn_norm / gcn_conv equal the textbook GCN  D^-1/2 (A + I) D^-1/2 X W^T + b  built densely with scipy
    (multi-edges counted with their multiplicity, existing self loops replaced by one, as add_remaining_self_loops
    does with unit weights), and hier_conv equals a dense row-normalised adjacency product."""
    import scipy.sparse as sp
    from oracle import tip_oracle as to
    rng = np.random.default_rng(11)
    n, e, fi, fo = 70, 500, 12, 7
    ei = rng.integers(0, n, size=(2, e))
    ei[:, :5] = np.stack([np.arange(5), np.arange(5)])          # a few self loops
    ei = np.concatenate([ei, ei[:, :40]], axis=1)                # and duplicated edges
    x = rng.standard_normal((n, fi))
    w = rng.standard_normal((fo, fi))
    b = rng.standard_normal(fo)
    got = to.gcn_conv(torch.from_numpy(x), to.gcn_norm(torch.from_numpy(ei), n, torch.float64), torch.from_numpy(w),
                      torch.from_numpy(b)).numpy()
    keep = ei[0] != ei[1]
    a = sp.coo_matrix((np.ones(int(keep.sum())), (ei[1][keep], ei[0][keep])), shape=(n, n)).tocsr() + sp.identity(n)
    dis = 1.0 / np.sqrt(np.asarray(a.sum(axis=1)).ravel())
    want = (sp.diags(dis) @ a @ sp.diags(dis)) @ (x @ w.T) + b
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)

    n_src, n_tgt, fh = 50, 9, 5
    dp = np.stack([rng.integers(0, n_src, size=120), n_src + rng.integers(0, n_tgt - 1, size=120)])   # last target: no edge
    xs = rng.standard_normal((n_src + n_tgt, fi))
    wh = rng.standard_normal((fi, fh))
    got = to.hier_conv(torch.from_numpy(xs), torch.from_numpy(dp), torch.from_numpy(wh), n_src, n_tgt).numpy()
    adj = sp.coo_matrix((np.ones(120), (dp[1], dp[0])), shape=(n_src + n_tgt, n_src + n_tgt)).toarray()
    deg = np.maximum(adj.sum(axis=1, keepdims=True), 1.0)
    want = ((adj / deg) @ xs)[n_src:] @ wh
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert np.all(got[-1] == 0.0)
