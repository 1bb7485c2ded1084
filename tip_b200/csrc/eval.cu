// Per-relation AUPRC / AUROC / AP on the GPU.
//
// Replaces TIP.compute_auprc_auroc_ap_by_et (src/layers.py:353-375), which for each of the 861 relations moves
// the scores to the host and calls sklearn three times through auprc_auroc_ap (src/utils.py:86-93):
//     y = [1]*k + [0]*k, pred = [pos_score[start:end], neg_score[start:end]]
//     auroc = roc_auc_score(y, pred);  ap = average_precision_score(y, pred)
//     p, r, _ = precision_recall_curve(y, pred);  auprc = auc(r, p)
// All three are sums over the DISTINCT score thresholds in decreasing order.  With tp_g / fp_g the positives /
// negatives at or above threshold g, P / N the class sizes, R_g = tp_g / P, Pr_g = tp_g / (tp_g + fp_g) and the
// start point (R_0, Pr_0) = (0, 1):
//     auroc = sum_g (fp_g - fp_{g-1}) (tp_g + tp_{g-1}) / 2 / (P N)          trapezoid under the ROC curve
//     ap    = sum_g (R_g - R_{g-1}) Pr_g
//     auprc = sum_g (R_g - R_{g-1}) (Pr_g + Pr_{g-1}) / 2                     trapezoid under the PR curve
// Program: two stable radix sorts (score descending, then relation) put every relation's 2k scores in order in one
// contiguous range; a global scan of the labels gives tp at every position; one CTA per relation finds the ends of
// the tie groups and accumulates the three sums in double precision in a fixed order.
#include "common.cuh"

namespace tipb {

constexpr int EVAL_T = 256;

__device__ __forceinline__ uint32_t desc_key(float v) {
    uint32_t b = __float_as_uint(v);
    b ^= (b & 0x80000000u) ? 0xffffffffu : 0x80000000u;  // ascending unsigned order == ascending float order
    return ~b;                                            // descending
}

// element id i in [0, 2E): i < E -> positive edge i, else negative edge i - E
__global__ void k_eval_score_keys(const float* __restrict__ pos, const float* __restrict__ neg, int64_t n_edges,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= 2 * n_edges) return;
    const float v = i < n_edges ? pos[i] : neg[i - n_edges];
    keys[i] = desc_key(v);
    vals[i] = uint32_t(i);
}

__global__ void k_eval_rel_keys(const uint32_t* __restrict__ vals, const int64_t* __restrict__ range_list,
                                int64_t n_edges, int n_rel, uint32_t* __restrict__ keys) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= 2 * n_edges) return;
    const int64_t e = vals[i] < n_edges ? vals[i] : vals[i] - n_edges;
    int lo = 0, hi = n_rel - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (range_list[2 * mid] <= e) lo = mid; else hi = mid - 1;
    }
    // an edge outside every range goes to an overflow bucket after the last relation
    keys[i] = (range_list[2 * lo] <= e && e < range_list[2 * lo + 1]) ? uint32_t(lo) : uint32_t(n_rel);
}

__global__ void k_eval_labels(const uint32_t* __restrict__ vals, int64_t n_edges, int* __restrict__ label) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < 2 * n_edges) label[i] = vals[i] < n_edges ? 1 : 0;
}

// tpx[i] = positives among sorted positions [0, i) (exclusive scan of the labels, tpx[2E] = total)
__global__ void __launch_bounds__(EVAL_T)
k_eval_metrics(const uint32_t* __restrict__ vals, const int* __restrict__ tpx, const float* __restrict__ pos,
               const float* __restrict__ neg, const int64_t* __restrict__ range_list, int64_t n_edges,
               int n_rel, double* __restrict__ record) {
    __shared__ int s_start[EVAL_T];
    __shared__ double s_red[3][EVAL_T];
    __shared__ int s_carry;
    const int r = blockIdx.x, t = threadIdx.x;
    const int64_t lo = 2 * range_list[2 * r], hi = 2 * range_list[2 * r + 1];
    const int64_t n = hi - lo;
    const int base_tp = n > 0 ? tpx[lo] : 0;
    const double P = n > 0 ? double(tpx[hi] - base_tp) : 0.0;
    const double N = double(n) - P;
    auto score = [&](int64_t i) -> float {
        const uint32_t v = vals[i];
        return v < n_edges ? pos[v] : neg[v - n_edges];
    };
    double a_roc = 0.0, a_ap = 0.0, a_pr = 0.0;
    if (t == 0) s_carry = 0;  // start (relative position) of the tie group running into the current tile
    __syncthreads();
    for (int64_t tile = 0; tile < n; tile += EVAL_T) {
        const int64_t i = tile + t;
        const bool in = i < n;
        float mine = 0.f, prev = 0.f, next = 0.f;
        if (in) {
            mine = score(lo + i);
            if (i > 0) prev = score(lo + i - 1);
            if (i + 1 < n) next = score(lo + i + 1);
        }
        const bool is_start = in && (i == 0 || prev != mine);
        const bool is_end = in && (i + 1 == n || next != mine);
        // inclusive max-scan of group starts over the tile (Hillis-Steele in shared memory, fixed order)
        s_start[t] = is_start ? int(i) : -1;
        __syncthreads();
        for (int o = 1; o < EVAL_T; o <<= 1) {
            const int other = t >= o ? s_start[t - o] : -1;
            __syncthreads();
            if (other > s_start[t]) s_start[t] = other;
            __syncthreads();
        }
        const int start = s_start[t] >= 0 ? s_start[t] : s_carry;
        if (is_end) {
            const double tp = double(tpx[lo + i + 1] - base_tp), fp = double(i + 1) - tp;
            const double tp0 = double(tpx[lo + start] - base_tp), fp0 = double(start) - tp0;
            const double pr = tp / (tp + fp), pr0 = start > 0 ? tp0 / (tp0 + fp0) : 1.0;
            const double dr = P > 0.0 ? (tp - tp0) / P : 0.0;
            a_roc += (fp - fp0) * (tp + tp0) * 0.5;
            a_ap += dr * pr;
            a_pr += dr * (pr + pr0) * 0.5;
        }
        __syncthreads();
        if (t == EVAL_T - 1) s_carry = start;  // the last lane's group start carries over (it is `in` unless this is the last tile)
        __syncthreads();
    }
    s_red[0][t] = a_pr; s_red[1][t] = a_roc; s_red[2][t] = a_ap;
    __syncthreads();
    for (int o = EVAL_T / 2; o > 0; o >>= 1) {
        if (t < o) {
            s_red[0][t] += s_red[0][t + o]; s_red[1][t] += s_red[1][t + o]; s_red[2][t] += s_red[2][t + o];
        }
        __syncthreads();
    }
    if (t == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        const bool ok = P > 0.0 && N > 0.0;  // sklearn raises when only one class is present
        record[r] = ok ? s_red[0][0] : nan;                        // auprc
        record[n_rel + r] = ok ? s_red[1][0] / (P * N) : nan;      // auroc
        record[2 * n_rel + r] = ok ? s_red[2][0] : nan;            // ap
    }
}

static int eval_bits_for(uint64_t v) {
    int b = 1;
    while (b < 32 && (uint64_t(1) << b) <= v) ++b;
    return b;
}

struct EvalWs { uint32_t *k0, *v0, *k1, *v1; int* tp; void *scan_ws, *sort_ws; };
static size_t eval_ws_layout(int64_t n_edges, void* ws, EvalWs* out) {
    const int64_t n = 2 * n_edges > 0 ? 2 * n_edges : 1;
    Carver c(ws);
    EvalWs w;
    w.k0 = c.take<uint32_t>(n); w.v0 = c.take<uint32_t>(n); w.k1 = c.take<uint32_t>(n); w.v1 = c.take<uint32_t>(n);
    w.tp = c.take<int>(n + 1);
    w.scan_ws = c.take<char>(scan_ws_bytes(n));
    w.sort_ws = c.take<char>(sort_ws_bytes(n));
    if (out) *out = w;
    return c.used() + 256;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

size_t tipb_eval_workspace_bytes(int64_t n_edges, int64_t n_rel) {
    (void)n_rel;
    return eval_ws_layout(n_edges, nullptr, nullptr);
}

int tipb_eval_auprc_auroc_ap(const float* pos_score, const float* neg_score, const int64_t* range_list, int64_t n_edges,
                             int64_t n_rel, double* record, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(range_list && record && ws && (n_edges == 0 || (pos_score && neg_score)), "eval: NULL argument");
    TIPB_CHECK_ARG(n_rel > 0 && n_edges >= 0 && 2 * n_edges < (int64_t(1) << 31) - 4096, "eval: sizes out of range");
    TIPB_CHECK_ARG(ws_bytes >= eval_ws_layout(n_edges, nullptr, nullptr), "eval: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    EvalWs w;
    eval_ws_layout(n_edges, ws, &w);
    const int64_t n = 2 * n_edges;
    int rc;
    if (n > 0) {
        const unsigned g = (unsigned)ceil_div(n, 256);
        k_eval_score_keys<<<g, 256, 0, s>>>(pos_score, neg_score, n_edges, w.k0, w.v0);
        if ((rc = sort_pairs_u32(w.k0, w.v0, w.k1, w.v1, n, 32, w.sort_ws, s))) return rc;
        k_eval_rel_keys<<<g, 256, 0, s>>>(w.v1, range_list, n_edges, (int)n_rel, w.k1);
        if ((rc = sort_pairs_u32(w.k1, w.v1, w.k0, w.v0, n, eval_bits_for(uint64_t(n_rel)), w.sort_ws, s))) return rc;
        k_eval_labels<<<g, 256, 0, s>>>(w.v0, n_edges, w.tp);
        if ((rc = exclusive_scan_i32(w.tp, w.tp, n, w.scan_ws, s))) return rc;
    } else {
        TIPB_CHECK_CUDA(cudaMemsetAsync(w.tp, 0, sizeof(int), s));
    }
    k_eval_metrics<<<(unsigned)n_rel, EVAL_T, 0, s>>>(w.v0, w.tp, pos_score, neg_score, range_list, n_edges, (int)n_rel,
                                                      record);
    TIPB_CHECK_LAUNCH("eval_auprc_auroc_ap");
    return TIPB_OK;
}
}
