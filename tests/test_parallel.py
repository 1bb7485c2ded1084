"""Relation sharding (tip_b200/parallel.py): host logic on CPU (gloo, world_size 2) and the N-rank == 1-rank
equivalence of the real CUDA path (two processes sharing cuda:0 through gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_relations_balanced_and_contiguous():
    from tip_b200.parallel import partition_relations
    rng = np.random.default_rng(0)
    sizes = np.clip(rng.lognormal(7.7, 1.2, 861), 250, 60000).astype(np.int64)
    ends = np.cumsum(sizes)
    rl = np.stack([ends - sizes, ends], axis=1)
    for world in (1, 2, 4, 8):
        blocks = partition_relations(rl, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == 861
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
        loads = np.array([sizes[a:b].sum() for a, b in blocks], dtype=np.float64)
        assert loads.sum() == sizes.sum()
        assert loads.max() / loads.mean() < 1.05, (world, loads)
    # degenerate inputs
    assert partition_relations(rl[:1], 4)[0] == (0, 1) or sum(b - a for a, b in partition_relations(rl[:1], 4)) == 1
    empty = np.zeros((3, 2), dtype=np.int64)
    assert sum(b - a for a, b in partition_relations(empty, 2)) == 3


def _cpu_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tip_b200.parallel import _Collective, _ReduceBwd, _ReduceFwd
    coll = _Collective(world)
    torch.manual_seed(0)
    x = torch.randn(5, 4, requires_grad=True)            # replicated input
    w_full = torch.randn(world, 4, 3)                    # one "relation" weight per rank
    root = torch.randn(4, 3, requires_grad=True)         # replicated parameter used on every rank
    w = w_full[rank].clone().requires_grad_(True)        # relation-local parameter
    part = _ReduceBwd.apply(x, coll) @ w + x.detach() @ (root / world)
    y = _ReduceFwd.apply(part, coll)
    loss = (y ** 2).sum()
    loss.backward()
    g_root = root.grad.clone()
    coll.all_reduce_(g_root)                             # what ShardedTIP.sync_gradients does
    # single-process reference
    x2 = x.detach().clone().requires_grad_(True)
    w2 = w_full.clone().requires_grad_(True)
    root2 = root.detach().clone().requires_grad_(True)
    y2 = sum(x2 @ w2[k] for k in range(world)) + x2.detach() @ root2
    ((y2 ** 2).sum()).backward()
    ok = (torch.allclose(y, y2, atol=1e-5) and torch.allclose(x.grad, x2.grad, atol=1e-4)
          and torch.allclose(w.grad, w2.grad[rank], atol=1e-4) and torch.allclose(g_root, root2.grad, atol=1e-4))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_collective_autograd_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cpu_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


# ------------------------------------------------------------------------------------------- GPU
def _make_small_data():
    from tip_b200 import synth
    return synth.make_tip_data(n_drug=120, n_prot=600, n_rel=23, dd_undirected=12_000, pp_undirected=3_000,
                               pd_edges=500, seed=5)


def _bench_data():
    import bench
    return bench.make_data("polypharmacy")[0]       # exactly the instance bench.py times


def _gpu_worker(rank, world, port, out, big=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tip_b200 import layers, neg_sampling as ns, parallel
    dev = torch.device("cuda:0")
    data = _bench_data() if big else _make_small_data()
    torch.manual_seed(1111)
    ns.seed(1111, dev)
    settings = layers.Setting(sp_rate=0.9, lr=0.01, prot_drug_dim=16, n_embed=48, n_hid1=32, n_hid2=16, num_base=32)
    model = parallel.ShardedTIP(settings, dev, mod="cat", data=data, rank=rank, world=world,
                                defer_loss_reduce=(rank >= 0 and big))     # both flavours of the loss exchange
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    losses = []
    for it in range(2):
        if it == 1 and not big:
            # this rank's edges are overwritten in place (what bench.py's end-to-end loop does every step): the shard's
            # index structures must be rebuilt and give the same numbers
            model.local_idx.copy_(model.local_idx.clone())
            model.refresh_shard()
        opt.zero_grad()
        loss = model()
        loss.backward()
        opt.step()
        losses.append(float(model.last_loss) if model.defer_loss_reduce else float(loss.detach()))
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters()}
    neg = model._neg_index.cpu()
    model.gather_parameters(opt)                      # every rank now holds ALL rows of att / decoder.weight
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    moments = {n: opt.state[p]["exp_avg"].detach().cpu() for n, p in model.named_parameters()}
    out[rank] = dict(losses=losses, grads=grads, z=model.embeddings.detach().cpu(), block=(model.r_lo, model.r_hi),
                     edges=(model.e_lo, model.e_hi), neg=neg, params=params, moments=moments,
                     test_neg=model.test_neg_index.cpu())
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("big", [False, True], ids=["small", "benched-workload"])
def test_two_rank_sharding_equals_single_rank(big):
    """`big`: the exact instance bench.py times (861 relations, 8.28 M edges) split over two ranks"""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tip_b200 import layers, neg_sampling as ns
    dev = torch.device("cuda:0")
    data = _bench_data() if big else _make_small_data()
    torch.manual_seed(1111)
    ns.seed(1111, dev)
    settings = layers.Setting(sp_rate=0.9, lr=0.01, prot_drug_dim=16, n_embed=48, n_hid1=32, n_hid2=16, num_base=32)
    ref = layers.TIP(settings, dev, mod="cat", data=data)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01)
    ref_losses = []
    for _ in range(2):
        opt.zero_grad()
        loss = ref()
        loss.backward()
        opt.step()
        ref_losses.append(float(loss.detach()))
    ref_grads = {n: p.grad.detach().cpu() for n, p in ref.named_parameters()}
    ref_neg = ref._neg_index.cpu()
    ref_z = ref.embeddings.detach().cpu()
    ref_params = {n: p.detach().cpu() for n, p in ref.named_parameters()}
    ref_moments = {n: opt.state[p]["exp_avg"].detach().cpu() for n, p in ref.named_parameters()}

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), out, big), nprocs=world, join=True)
    res = dict(out)
    assert set(res) == {0, 1}
    for rank in (0, 1):
        r = res[rank]
        e_lo, e_hi = r["edges"]
        assert torch.equal(r["neg"], ref_neg[:, e_lo:e_hi]), "negative pairs must not depend on the sharding"
        assert torch.equal(r["test_neg"], ref.test_neg_index.cpu()), "test-set negatives (drawn first from the same stream)"
        np.testing.assert_allclose(r["losses"], ref_losses, rtol=2e-5)
        torch.testing.assert_close(r["z"], ref_z, rtol=1e-4, atol=1e-5)
        lo, hi = r["block"]
        for n, g in r["grads"].items():
            want = ref_grads[n]
            if n.endswith("rgcn1.att") or n.endswith("rgcn2.att") or n == "decoder.weight":
                g, want = g[lo:hi], want[lo:hi]           # relation-local rows live on their owner
            scale = float(want.abs().max()) + 1e-30
            torch.testing.assert_close(g, want, rtol=1e-3, atol=5e-6 * scale, msg=lambda m: f"rank {rank} {n}: {m}")
        # after gather_parameters every rank holds the single-GPU model (rows of other ranks were stale)
        for n, p in r["params"].items():
            scale = float(ref_params[n].abs().max()) + 1e-30
            torch.testing.assert_close(p, ref_params[n], rtol=2e-3, atol=2e-5 * scale, msg=lambda m: f"rank {rank} param {n}: {m}")
            scale = float(ref_moments[n].abs().max()) + 1e-30
            torch.testing.assert_close(r["moments"][n], ref_moments[n], rtol=2e-3, atol=2e-5 * scale,
                                       msg=lambda m: f"rank {rank} exp_avg {n}: {m}")
    for n in res[0]["params"]:
        assert torch.equal(res[0]["params"][n], res[1]["params"][n]), "replicas must be bit-identical after the gather: " + n
