"""ctypes binding of libtipb200.so (the C ABI declared in include/tipb200.h).

There is no fallback: if the shared library is missing or a call fails, this
module raises.  Pointers are passed as raw device addresses (`tensor.data_ptr()`)
and the stream is torch's current CUDA stream, so calls are CUDA-graph capturable.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtipb200.so")

_p, _i64, _i32, _sz, _u32 = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_uint32

# name -> (restype, argtypes); mirrors include/tipb200.h one to one
SIGNATURES = {
    "tipb_version": (C.c_int, []),
    "tipb_last_error": (C.c_char_p, []),
    "tipb_scan_workspace_bytes": (_sz, [_i64]),
    "tipb_exclusive_scan_i32": (C.c_int, [_p, _p, _i64, _p, _sz, _p]),
    "tipb_sort_workspace_bytes": (_sz, [_i64]),
    "tipb_sort_pairs_u32": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _sz, _p]),
    "tipb_typed_csr_bytes": (_sz, [_i64, _i64, _i64]),
    "tipb_typed_csr_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "tipb_typed_csr_layout": (C.c_int, [_i64, _i64, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "tipb_typed_csr_build": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _p, _sz, _p, _sz, _p]),
    "tipb_seg_aggregate": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _i32, _p, _p]),
    "tipb_rgcn_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32, _i32]),
    "tipb_rgcn_fwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "tipb_rgcn_tc_status": (C.c_int, []),
    "tipb_rgcn_bwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32,
                                _p, _p, _p, _p, _p, _p, _sz, _p]),
    "tipb_gcn_norm": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "tipb_gcn_spmm": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p, _i32, _i32, _p, _p]),
    "tipb_hier_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "tipb_hier_fwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i32, _i32, _p, _p, _p]),
    "tipb_hier_bwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _sz, _p]),
    "tipb_decoder_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _p, _p]),
    "tipb_decoder_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "tipb_decoder_bwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i32, _i32, _p, _p, _p, _sz, _p]),
    "tipb_decoder_bce_fused": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "tipb_decoder_bce_fused_mirrored": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "tipb_edges_mirrored": (C.c_int, [_p, _p, _i64, _i64, _p, _p]),
    "tipb_decoder_sweep": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _p, _p]),
    "tipb_decoder_sweep_status": (C.c_int, []),
    "tipb_neg_bitmap_bytes": (_sz, [_i64, _i64]),
    "tipb_neg_bitmap_build": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "tipb_neg_table_build": (C.c_int, [_p, _p, _i64, _i64, C.c_double, _p, _p]),
    "tipb_neg_sample_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64, _i64]),
    "tipb_neg_sample": (C.c_int, [_p, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "tipb_neg_bitmap_build_range": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "tipb_neg_sample_shard_begin": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p,
                                              _sz, _p]),
    "tipb_neg_sample_shard_end": (C.c_int, [_p, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _i32, _i32, _i64,
                                            _i64, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "tipb_pair_pass_supported": (C.c_int, [_i64, _i32]),
    "tipb_pair_chunk": (_i64, []),
    "tipb_pair_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "tipb_pack_half_pairs": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "tipb_unpack_pairs": (C.c_int, [_p, _i64, _p, _p]),
    "tipb_pair_bce_pass": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _i32, C.c_float, C.c_float, _p, _sz, _p]),
    "tipb_pair_bce_finish": (C.c_int, [_p, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "tipb_eval_workspace_bytes": (_sz, [_i64, _i64]),
    "tipb_eval_auprc_auroc_ap": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _sz, _p]),
    "tipb_adam_max_tensors": (C.c_int, []),
    "tipb_adam_step": (C.c_int, [C.c_int, _p, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, C.c_double, _p, _p]),
    "tipb_gemm": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i32, _i32, _p, _p]),
    "tipb_gemm_tn_workspace_bytes": (_sz, [_i64, _i64]),
    "tipb_gemm_tn": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _sz, _p]),
    "tipb_relu_grad_colsum_workspace_bytes": (_sz, [_i64]),
    "tipb_relu_grad_colsum": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "tipb_transpose": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "tipb_drug_input_fwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "tipb_drug_input_bwd": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "tipb_scale2": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _p]),
    "tipb_fill_zero": (C.c_int, [_p, _sz, _p]),
    "tipb_nn_decoder_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i32, _p, _p]),
    "tipb_nn_decoder_bwd": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _p, _p, _p, _p]),
    "tipb_spmm_values": (C.c_int, [_p, _i64, _i64, _p, _p, _i32, _p, _p]),
    "tipb_edge_split_workspace_bytes": (_sz, [_i64]),
    "tipb_edge_split_mask": (C.c_int, [_p, _p, _i64, _i64, C.c_double, C.c_double, _i32, _p, _p, _p, _sz, _p]),
    "tipb_edge_split_emit": (C.c_int, [_p, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "tipb_mt19937_seed": (C.c_int, [_p, _u32, _p]),
    "tipb_mt19937_stream_words": (_i64, [_i64]),
    "tipb_mt19937_generate": (C.c_int, [_p, _p, _i64, _p]),
    "tipb_mt19937_jump_polys": (C.c_int, [_i64, _i64, _p]),
    "tipb_mt19937_chunk_count": (_i64, [_i64, _i64]),
    "tipb_mt19937_generate_chunked": (C.c_int, [_p, _p, _i64, _i64, _p, _i64, _p, _p]),
}

CSR_FIELDS = ("counts", "eid", "other", "seg_ptr", "seg_node", "seg_rel", "node_ptr", "deg", "inv_deg",
              "rel_seg_ptr", "rel_seg")

_lib = None


class TipbError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TipbError(
                f"{LIB_PATH} not found: build it with `make -C tip_b200/csrc` (or __graft_entry__.build()). "
                "tip_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise TipbError(f"{what} failed (code {rc}): {lib().tipb_last_error().decode()}")


def ptr(t):
    """device address of a tensor (None -> NULL); requires a contiguous CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TipbError("tip_b200 operators take CUDA tensors only (there is no CPU path)")
    if not t.is_contiguous():
        raise TipbError("internal error: non-contiguous tensor reached the C ABI")
    return t.data_ptr()


def stream(device=None):
    """raw handle of torch's current stream on `device` (default: the current device)"""
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t):
    """context: make the device of tensor/device `t` current.  The library launches on the CURRENT device (its
    pointers must live there), so every public entry point of tip_b200 wraps its calls in this guard -- the reference
    API selects the GPU through tensors / the `device` argument, not through torch.cuda.set_device."""
    dev = t.device if torch.is_tensor(t) else torch.device(t)
    if dev.type != "cuda":
        raise TipbError("tip_b200 operators take CUDA tensors only (there is no CPU path)")
    return torch.cuda.device(dev)


def csr_layout(n_entries, n_nodes, n_rel):
    offs = (_i64 * len(CSR_FIELDS))()
    cap = _i64()
    check(lib().tipb_typed_csr_layout(n_entries, n_nodes, n_rel, offs, C.byref(cap)), "typed_csr_layout")
    return dict(zip(CSR_FIELDS, list(offs))), cap.value
