"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/tipb200.h declares, the ctypes table matches the header, size queries work without a GPU,
and the product refuses CPU tensors instead of falling back."""
import os
import re

import pytest
import torch

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "tipb200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tipb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tip_b200 import _lib
    handle = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(handle, n), f"{n} declared in tipb200.h but not exported by libtipb200.so"
        assert n in _lib.SIGNATURES, f"{n} declared in tipb200.h but missing from the ctypes table"
    assert set(_lib.SIGNATURES) == set(names)
    assert handle.tipb_version() == 100


def test_size_queries_need_no_gpu():
    from tip_b200 import _lib
    L = _lib.lib()
    e, n, r = 8_280_000, 645, 861
    assert L.tipb_typed_csr_bytes(e, n, r) > 4 * 2 * e
    assert L.tipb_typed_csr_workspace_bytes(e, n, r) > 0
    layout, cap = _lib.csr_layout(e, n, r)
    assert cap == n * r and layout["counts"] == 0 and all(v % 256 == 0 for v in layout.values())
    assert L.tipb_rgcn_workspace_bytes(e, n, r, 64, 32, 32) >= n * r * 64 * 4
    assert L.tipb_neg_sample_workspace_bytes(e, r, 13_000_000, 11_000_000, 2_000_000) > 13_000_000 * 12
    assert L.tipb_mt19937_stream_words(1000) == 1362 and L.tipb_mt19937_stream_words(908) == 908
    assert L.tipb_neg_bitmap_bytes(n, r) == ((n * n + 31) // 32) * 4 * r
    assert L.tipb_sort_workspace_bytes(e) > 0 and L.tipb_scan_workspace_bytes(e) > 0


def test_neg_table_build_is_host_only_and_consistent():
    import ctypes as C

    import numpy as np
    from tip_b200 import _lib
    L = _lib.lib()
    sizes = np.array([1500, 0, 2500, 40, 9000], dtype=np.int64)
    ends = np.cumsum(sizes)
    rl = np.ascontiguousarray(np.stack([ends - sizes, ends], axis=1))
    pop = np.array([1400, 0, 2300, 40, 8000], dtype=np.int32)
    table = np.zeros((5, 8), dtype=np.int64)
    totals = np.zeros(4, dtype=np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(L.tipb_neg_table_build(vp(rl), vp(pop), 5, 101, 7.0, vp(table), vp(totals)), "table")
    lo, W, Lw, win_off, f_off, k, pred, order = table.T
    assert sorted(order) == list(range(5)) and np.all(np.diff(Lw[order]) <= 0)      # longest window first
    assert np.array_equal(k, sizes)
    assert lo[0] == 0 and np.all(np.diff(lo) >= 0)
    assert np.all(W >= 1) and np.all((Lw >= W + k) | (k == 0))
    Lp = (Lw + 3) & ~3                                    # window rows start 16-byte aligned (vector stores)
    assert np.array_equal(win_off, np.concatenate([[0], np.cumsum(Lp)[:-1]])) and np.all(win_off % 4 == 0)
    assert np.array_equal(f_off, np.concatenate([[0], np.cumsum(W)[:-1]]))
    assert totals[0] == Lp.sum() and totals[1] == W.sum() and totals[2] >= (lo + Lw).max()
    # expected offsets sit inside the brackets: relation r starts near sum(k) + sum(k d/(1-d))
    d = pop / 101.0 ** 2
    mean = np.concatenate([[0], np.cumsum(sizes * d / (1 - d))[:-1]])
    start = np.concatenate([[0], ends[:-1]]) + mean
    assert np.all(lo <= start) and np.all(start <= lo + W)


def test_mt19937_jump_polynomials_against_the_oracle_stream():
    """tipb_mt19937_jump_polys (host only): g_k(T) applied to a key block by plain Horner == the window the
    oracle's MT19937 reaches after k*chunk words (except the 31 unused low bits of word 0)."""
    import ctypes as C

    import numpy as np
    from oracle import neg_sampling_oracle as nso
    from tip_b200 import _lib
    L = _lib.lib()
    chunk, n_polys = 4540, 2
    polys = np.zeros((n_polys, 624), dtype=np.uint32)
    _lib.check(L.tipb_mt19937_jump_polys(chunk, n_polys, polys.ctypes.data_as(C.c_void_p)), "jump_polys")
    assert L.tipb_mt19937_chunk_count(10 * chunk + 454, chunk) == 11
    mt = nso.MT19937(1111)
    w0 = np.asarray(mt.key, dtype=np.uint32).copy()

    def advance(w):      # the one-word transition T on a 624-word window
        y = (w[0] & np.uint32(0x80000000)) | (w[1] & np.uint32(0x7fffffff))
        v = w[397] ^ (y >> np.uint32(1)) ^ (np.uint32(0x9908b0df) if (w[1] & np.uint32(1)) else np.uint32(0))
        out = np.empty_like(w)
        out[:-1] = w[1:]
        out[-1] = v
        return out

    want, w = {}, w0.copy()
    for n in range(1, n_polys * chunk + 1):
        w = advance(w)
        if n % chunk == 0:
            want[n // chunk] = w.copy()
    # the oracle's own generator agrees with `advance` (its key after one refill = the window 624 words on)
    mt._twist()
    w624 = w0.copy()
    for _ in range(624):
        w624 = advance(w624)
    assert np.array_equal(np.asarray(mt.key, dtype=np.uint32), w624)
    for k in range(1, n_polys + 1):
        bits = np.unpackbits(polys[k - 1].view(np.uint8), bitorder="little")
        assert not bits[19937:].any()
        h = np.zeros(624, dtype=np.uint32)
        for i in range(19936, -1, -1):
            h = advance(h)
            if bits[i]:
                h ^= w0
        assert np.array_equal(h[1:], want[k][1:]) and (h[0] >> 31) == (want[k][0] >> 31), k


def test_argument_errors_are_reported_not_thrown():
    from tip_b200 import _lib
    L = _lib.lib()
    rc = L.tipb_typed_csr_build(None, None, None, 5, 10, 10, 3, 0, 0, 0, 0, None, 0, None, 0, None)
    assert rc == -1 and b"typed_csr_build" in L.tipb_last_error()


def test_no_cpu_fallback():
    from tip_b200 import TipbError, layers
    conv = layers.MyRGCNConv2(8, 4, 3, 2, after_relu=False)
    ei = torch.zeros((2, 4), dtype=torch.long)
    with pytest.raises(TipbError):
        conv(torch.randn(5, 8), ei, torch.zeros(4, dtype=torch.long), torch.tensor([[0, 4], [4, 4], [4, 4]]))
    dec = layers.MultiInnerProductDecoder(4, 3)
    with pytest.raises(TipbError):
        dec(torch.randn(5, 4), ei, torch.zeros(4, dtype=torch.long))
    with pytest.raises(TipbError):
        layers.typed_negative_sampling(ei, 5, torch.tensor([[0, 4]]))
    # the later additions refuse host tensors too: evaluation, mirrored-layout check, sweep, optimiser
    from tip_b200 import ops, optim
    with pytest.raises(TipbError):
        ops.eval_auprc_auroc_ap(torch.rand(4), torch.rand(4), torch.tensor([[0, 4]]))
    with pytest.raises(TipbError):
        ops.edges_mirrored(ei, torch.tensor([[0, 4]]))
    with pytest.raises(TipbError):
        ops.decoder_sweep(torch.randn(5, 4), torch.randn(3, 4))
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.ones(3)
    with pytest.raises(TipbError):
        optim.Adam([p], lr=0.1).step()
    with pytest.raises(ValueError):
        optim.Adam([p], lr=-1.0)


def test_range_table_validation():
    from tip_b200 import ops
    ops.check_cumulative_ranges(torch.tensor([[0, 3], [3, 3], [3, 9]]), 9)
    for bad, e in ((torch.tensor([[0, 3], [4, 9]]), 9), (torch.tensor([[0, 3], [3, 8]]), 9), (torch.tensor([[1, 9]]), 9),
                   (torch.tensor([[0, 5], [5, 4]]), 4)):
        with pytest.raises(ValueError):
            ops.check_cumulative_ranges(bad, e)


def test_adam_argument_checks_need_no_gpu():
    """tipb_adam_step validates its host-side arguments before any launch"""
    import ctypes as C
    from tip_b200 import _lib
    L = _lib.lib()
    assert L.tipb_adam_max_tensors() >= 13          # the 13 parameter tensors of TIP in one launch
    rc = L.tipb_adam_step(1, None, None, None, None, None, 0.01, 0.9, 0.999, 1e-8, None, None)
    assert rc == -1 and b"adam_step" in L.tipb_last_error()
    one = (C.c_void_p * 1)(8)
    n = (C.c_int64 * 1)(4)
    rc = L.tipb_adam_step(1, one, one, one, one, n, 0.01, 1.5, 0.999, 1e-8, one, None)
    assert rc == -1 and b"hyper-parameter" in L.tipb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_init_draw_order_matches_reference_modules(golden_layers):
    """same seed -> same parameters as the reference's own modules (values from tests/golden)."""
    from tip_b200 import layers
    torch.manual_seed(1111)
    x = torch.randn(40, 24)   # make_golden.py draws x before building the conv
    conv = layers.MyRGCNConv2(24, 12, 6, 5, after_relu=False)
    assert torch.equal(x, torch.from_numpy(golden_layers["rgcn2/x"]))
    for name in ("att", "root", "basis"):
        assert torch.equal(getattr(conv, name).data, torch.from_numpy(golden_layers[f"rgcn2/{name}"])), name


def test_adam_state_layout_is_torchs():
    """the step counter must live in state[p] (as torch.optim.Adam keeps it), never in param_groups, and a
    loaded state_dict must make the optimiser re-derive its device counter (checked on the GPU in
    test_fused_adam_matches_torch_adam; here: the host-side bookkeeping, no launch)."""
    import torch
    from tip_b200 import optim
    from tip_b200._lib import TipbError
    p = [torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(2, 2))]
    ref = torch.optim.Adam(p, lr=0.1)
    for q in p:
        q.grad = torch.ones_like(q)
    ref.step()
    ref.step()
    mine = optim.Adam(p, lr=0.1)
    mine._counters[0] = "stale"
    mine.load_state_dict(ref.state_dict())
    assert mine._counters == {}                                       # re-derived at the next step
    assert "step" not in mine.state_dict()["param_groups"][0]
    assert "step" not in optim.Adam(p, lr=0.1).state_dict()["param_groups"][0]
    assert all(float(mine.state[q]["step"]) == 2.0 for q in p)
    with pytest.raises(TipbError):                                    # CPU parameters: no CPU path
        mine.step()
