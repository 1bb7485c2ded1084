"""Algorithmic-byte model of the kernels of one TIP training step (figure (A)) and the grouping
of kernels into families.  Shared by bench.py (live roofline block) and tools/roofline_table.py (ncu tables).

(A) counts, per launch: the index stream + the gathered payload rows + the outputs -- payload rows are served from
shared memory on the drug graph, so (A)/t can exceed the HBM peak; it is always reported next to the DRAM traffic
that ncu measured for the same kernel (profiles/*_kernel_traffic.json)."""
import re

# workload dims W: E (directed D-D edges of this rank), S_dst / S_src (non-empty (node, relation) segments),
# R, N (drugs), Np (proteins), Epp (directed P-P edges incl. self loops), B (bases), words (MT19937 words per step),
# sum_l, sum_w (sampler window sizes), F0/F1/F2 (R-GCN widths), dim (decoder width)

FAMILIES = [
    ("P-P GCN + hierarchy", r"k_node_aggregate|k_hier_|k_gcn_|k_lin_|k_gemm|k_transpose|k_mask_colsum|k_colsum_finish|k_drug_input"),
    ("R-GCN layer 1 fwd", None), ("R-GCN layer 2 fwd", None), ("R-GCN layer 2 bwd", None), ("R-GCN layer 1 bwd", None),
    ("decoder + loss (pos+neg, fwd+bwd)", r"k_pair_|k_decoder_|k_loss_reduce|k_add_inplace|k_scale2"),
    ("negative sampler", r"k_accept_count|k_compact|k_window_scan|k_chain_|k_materialize|k_finalize|k_mt_"),
    ("negative plan build", r"k_grp_|k_csr_|k_rel_order|k_scan_lookback|k_sort_"),
    ("Adam", r"k_adam"),
    ("collectives (NCCL kernels, incl. waiting for the peers)", r"nccl|ncclDevKernel"),
]


def short_name(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("tipb::", "")
    name = name.replace("(bool)1", "true").replace("(bool)0", "false")      # ncu prints template bools as (bool)1
    return re.sub(r"\(int\)|\(bool\)", "", name).strip()


class StepModel(object):
    """walks the kernels of ONE serial step in launch order and labels each with its family and (A) bytes"""

    def __init__(self, W):
        self.W = dict(W)
        self.rgcn_phase = 0          # seg_aggregate launches seen: 1 -> L1 fwd, 2 -> L2 fwd, 3 -> L2 bwd, 4 -> L1 bwd
        self.seen = {}

    def _rgcn_family(self):
        return ["P-P GCN + hierarchy", "R-GCN layer 1 fwd", "R-GCN layer 2 fwd", "R-GCN layer 2 bwd",
                "R-GCN layer 1 bwd", "P-P GCN + hierarchy"][min(self.rgcn_phase, 5)]

    def label(self, name):
        """-> (family, algorithmic bytes or None, formula or None)"""
        W = self.W
        s = short_name(name)
        E, S, Ssrc, R, N, B = W["E"], W["S_dst"], W["S_src"], W["R"], W["N"], W.get("B", 32)
        k = self.seen.get(s, 0)
        self.seen[s] = k + 1
        if "k_hier_bwd" in s:        # the D-D backward is over: what follows is the P-D / P-P backward
            self.rgcn_phase = 5
        if "k_rgcn_grad_prep" in s and self.rgcn_phase in (2, 3):     # opens the next backward layer
            return ("R-GCN layer 2 bwd" if self.rgcn_phase == 2 else "R-GCN layer 1 bwd"), None, None
        m = re.match(r"k_seg_aggregate(?:_flat)?<(\d+)", s)
        if m:
            self.rgcn_phase += 1
            f = int(m.group(1)) * (1 if "flat" in s else 4)
            seg = S if self.rgcn_phase <= 2 else Ssrc
            return self._rgcn_family(), E * (4 + 4 * f) + seg * (4 * f + 4), "E(4+4F)+S(4F+4), F=%d" % f
        m = re.match(r"k_rgcn_node_fwd(?:_tiled)?<(\d+)", s)
        if m and "tiled" in s:
            f = int(m.group(1)) * 4
            return self._rgcn_family(), S * (4 * f + 4 * B + 12) + N * 4 * f, "S(4F_in+4B+12)+N*4F, F_in=%d" % f
        m = re.match(r"k_rgcn_node_bwd(?:_tiled)?<(\d+)", s)
        if m and "tiled" in s:
            f = int(m.group(1)) * 4
            return self._rgcn_family(), Ssrc * (4 * f + 8 * B + 12) + N * 8 * f, "S(4F_out+8B+12)+N*8F, F_out=%d" % f
        m = re.match(r"k_rgcn_node_fwd_tc<(\d+)", s)
        if m:
            f = int(m.group(1))
            return self._rgcn_family(), S * (4 * f + 4 * B + 4) + N * 4 * B * f, "S(4F_in+4B+4)+N*4BF, F_in=%d" % f
        m = re.match(r"k_rgcn_node_bwd_tc<(\d+)", s)
        if m:
            f = int(m.group(1))
            return self._rgcn_family(), Ssrc * (4 * f + 8 * B + 4) + N * 8 * B * f, "S(4F_out+8B+4)+N*8BF, F_out=%d" % f
        if re.match(r"k_rgcn_|k_atb|k_sum_slices|k_rel_reduce|k_rd_", s):
            return self._rgcn_family(), None, None
        m = re.match(r"k_pair_pass<(\d+)", s)
        if m:
            dim = int(m.group(1))
            # per scored entry 8 B of indices + two gathered rows; ONE launch scores the E positive and the
            # E negative entries (fwd + gradient fused)
            return FAMILIES[5][0], 2 * E * (8 + 2 * 4 * dim), "2E(8+2*4d), d=%d (SURVEY 8d decoder figure; pos + neg entries)" % dim
        m = re.match(r"k_decoder_seg<(\d+), (\d+)", s)
        if m:
            dim, mode = int(m.group(1)) * 4, int(m.group(2))
            return FAMILIES[5][0], E * (8 + 2 * 4 * dim), "E(8+2*4d), d=%d" % dim
        m = re.match(r"k_node_aggregate<(\d+)", s)
        if m:
            f = int(m.group(1)) * 4
            return FAMILIES[0][0], W["Epp"] * (4 + 4 + 4 * f) + W["Np"] * (4 + 4 * f), "E'pp(8+4F)+Np(4+4F), F=%d" % f
        if "k_materialize_main" in s:
            return FAMILIES[6][0], E * (4 + 4 + 4), "E(4 accepted + 4 bitmap + 4 packed out)"
        if "k_accept_count" in s:
            return FAMILIES[6][0], W["words"] * 4, "4 B per MT word"
        if "k_compact" in s:
            return FAMILIES[6][0], W["words"] * 4 + int(W["words"] * 0.79) * 8, "4/word + 8/accepted"
        if "k_window_scan" in s:
            return FAMILIES[6][0], W["sum_l"] * 8 + W["sum_w"] * 8, "8 sum_L + 8 sum_W"
        if "k_mt_generate_chunks" in s:
            return FAMILIES[6][0], W["words"] * 4, "4 B per word written"
        if "k_grp_place" in s:
            return FAMILIES[7][0], 2 * E * (16 + 8), "2E(16 read + 8 written)"
        for fam, pat in FAMILIES:
            if pat and re.search(pat, s):
                return fam, None, None
        return "other (torch / library kernels)", None, None


def workload_dims(model):
    """dims of the (rank's share of the) workload a model instance runs; one-time host reads"""
    from tip_b200 import neg_sampling as ns, ops
    d = model.data
    idx = getattr(model, "local_idx", d.dd_train_idx)
    rl = getattr(model, "local_range", d.dd_train_range)
    n_rel = max(int(rl.shape[0]), 1)
    p_dst = ops.cached_plan(idx, d.n_drug, n_rel, by_src=False, edge_type=None, range_list=rl)
    p_src = ops.cached_plan(idx, d.n_drug, n_rel, by_src=True, edge_type=None, range_list=rl)
    m = ns._membership(d.dd_train_idx, d.n_drug, d.dd_train_range)
    rng = ns._get_rng(model.device)
    enc = model.encoder
    return dict(E=int(idx.shape[1]), E_global=int(d.dd_train_idx.shape[1]), N=int(d.n_drug), Np=int(d.n_prot), R=n_rel,
                R_global=int(d.n_dd_et), S_dst=int(p_dst.field("counts")[0]), S_src=int(p_src.field("counts")[0]),
                Epp=int(d.pp_train_indices.shape[1]) + int(d.n_prot), E_pd=int(d.dp_edge_index.shape[1]),
                words=int(624 + rng.n_new), sum_l=int(m.sum_l), sum_w=int(m.sum_w), B=int(enc.rgcn1.num_bases),
                F0=int(enc.rgcn1.in_channels), F1=int(enc.rgcn1.out_channels), F2=int(enc.rgcn2.out_channels),
                dim=int(model.settings.n_hid2))


def step_algorithmic_bytes(W):
    """SURVEY 8(d) step-level figure: sum of (A) over the fwd+bwd edge passes, decoder, P-P, sampler"""
    E, F0, F1, F2, d = W["E"], W["F0"], W["F1"], W["F2"], W["dim"]
    rgcn_fwd = E * (4 + 4 * F0) + E * (4 + 4 * F1)
    rgcn_bwd = E * (4 + 4 * F1) + E * (4 + 4 * F0) + E * (4 + 4 * F2) + E * (4 + 4 * F1)
    dec = 2 * E * (8 + 2 * 4 * d)
    pp = 2 * (W["Epp"] * (8 + 4 * 32) + W["Epp"] * (8 + 4 * 16))
    sampler = E * 25
    return rgcn_fwd + rgcn_bwd + dec + pp + sampler
