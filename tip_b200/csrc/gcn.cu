// P-P GCN (symmetric-normalised CSR SpMM, north_star item 1) and the P->D hierarchy conv (item 2).
//
// GCNConv (torch_geometric 2.0.1, as called at reference src/layers.py:386-387, 391-395):
//   A' = A (self loops dropped) + I,  deg_i = in-degree_i + 1,  dis = deg^-1/2
//   out_i = sum_{j->i} dis_j dis_i x'_j + dis_i^2 x'_i + bias
// evaluated as out_i = dis_i * ( sum_j dis_j x'_j + dis_i x'_i ) + bias so that no per-edge weight is
// ever stored: the CSR carries indices only (4 B per edge).  One warp per destination row, each feature
// row read as 128-bit words by F/4 lanes (32/(F/4) neighbours in flight per instruction).
// The backward pass is the same kernel on the by-source CSR (A'^T), with the dis roles swapped.
//
// MyHierarchyConv (reference src/layers.py:229-242): mean over incoming edges, rows
// [n_source, n_source+n_target), then @ weight.  Same row kernel + a fused F_in x F_out epilogue.
#include "common.cuh"
#include "seg_aggregate.cuh"

namespace tipb {

// out[n] = post( scale_out[n] * ( sum_{p in node n} scale_in[other[p]] * x[other[p]] + self * scale_in[n] * x[n] ) )
// rows are nodes [node_lo, node_hi) of the plan; output row index is n - node_lo.
constexpr int HUB_DEGREE = 256;  // rows longer than this are split over the eight warps of their CTA

template <int LPR>
__device__ __forceinline__ void node_epilogue(float4 a, int n, int l, const float* __restrict__ scale_out,
                                              const float* __restrict__ scale_in, const float4* __restrict__ x,
                                              const float4* __restrict__ bias, int node_lo, int self_term, int relu,
                                              float4* __restrict__ out) {
    if (self_term) {
        const float si = scale_in ? scale_in[n] : 1.f;
        a = f4_fma(si, x[int64_t(n) * LPR + l], a);
    }
    if (scale_out) {
        const float so = scale_out[n];
        a.x *= so; a.y *= so; a.z *= so; a.w *= so;
    }
    if (bias) a = f4_add(a, bias[l]);
    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    out[int64_t(n - node_lo) * LPR + l] = a;
}

// Each CTA owns eight consecutive rows.  Short rows: one warp per row.  Hub rows (the P-P graph has proteins
// with thousands of neighbours) would serialise on one warp, so all eight warps take an eighth of the row each
// and the partial sums are added in warp order (fixed order => deterministic).
template <int LPR>
__global__ void __launch_bounds__(256)
k_node_aggregate(const int* __restrict__ node_ptr, const int* __restrict__ seg_ptr, const int* __restrict__ other,
                 const float* __restrict__ scale_out, const float* __restrict__ scale_in, const float4* __restrict__ x,
                 const float4* __restrict__ bias, int node_lo, int node_hi, int self_term, int relu,
                 float4* __restrict__ out) {
    __shared__ float4 part[8][LPR];
    __shared__ int s_beg[8], s_end[8];
    const int lane = lane_id(), w = warp_id();
    const int g = lane / LPR, l = lane % LPR;
    for (int n0 = node_lo + blockIdx.x * 8; n0 < node_hi; n0 += gridDim.x * 8) {
        const int n = n0 + w;
        int beg = 0, end = 0;
        if (n < node_hi) {
            beg = seg_ptr[node_ptr[n]];
            end = seg_ptr[node_ptr[n + 1]];
        }
        if (lane == 0) { s_beg[w] = beg; s_end[w] = end; }
        if (n < node_hi && end - beg <= HUB_DEGREE) {
            const int idx = (beg + lane < end) ? ld_stream_i32(other + beg + lane) : 0;
            const float4 a = warp_gather_sum<LPR, false>(x, other, scale_in, beg, end, idx);
            if (g == 0) node_epilogue<LPR>(a, n, l, scale_out, scale_in, x, bias, node_lo, self_term, relu, out);
        }
        __syncthreads();
        for (int h = 0; h < 8; ++h) {            // CTA-uniform: every warp sees the same row bounds
            const int hb = s_beg[h], he = s_end[h];
            if (he - hb <= HUB_DEGREE) continue;
            const int chunk = (he - hb + 7) >> 3;
            const int cb = min(hb + w * chunk, he), ce = min(cb + chunk, he);
            const int idx = (cb + lane < ce) ? ld_stream_i32(other + cb + lane) : 0;
            const float4 a = warp_gather_sum<LPR, false>(x, other, scale_in, cb, ce, idx);
            if (g == 0) part[w][l] = a;
            __syncthreads();
            if (w == 0 && g == 0) {
                float4 t = part[0][l];
#pragma unroll
                for (int k = 1; k < 8; ++k) t = f4_add(t, part[k][l]);
                node_epilogue<LPR>(t, n0 + h, l, scale_out, scale_in, x, bias, node_lo, self_term, relu, out);
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

static int node_aggregate_launch(const CsrView& v, const float* scale_out, const float* scale_in, const float* x,
                                 const float* bias, int node_lo, int node_hi, int f, int self_term, int relu,
                                 float* out, cudaStream_t s) {
    const int rows = node_hi - node_lo;
    if (rows <= 0) return TIPB_OK;
    int blocks = (int)ceil_div(rows, 8);
    const int cap = sm_count() * 16;
    if (blocks > cap) blocks = cap;
#define NA(LPRV)                                                                                                   \
    k_node_aggregate<LPRV><<<blocks, 256, 0, s>>>(v.node_ptr, v.seg_ptr, v.other, scale_out, scale_in, (const float4*)x, \
                                                  (const float4*)bias, node_lo, node_hi, self_term, relu, (float4*)out)
    switch (f) {
        case 4: NA(1); break;
        case 8: NA(2); break;
        case 16: NA(4); break;
        case 32: NA(8); break;
        case 64: NA(16); break;
        case 128: NA(32); break;
        default:
            set_last_error("node_aggregate: feature width %d not in {4,8,16,32,64,128}", f);
            return TIPB_ERR_UNSUPPORTED;
    }
#undef NA
    TIPB_CHECK_LAUNCH("node_aggregate");
    return TIPB_OK;
}

__global__ void k_gcn_norm(const int* __restrict__ deg, int n, float* __restrict__ dis) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dis[i] = 1.0f / sqrtf(float(deg[i] + 1));
}

// hierarchy forward epilogue: mean[t] = inv_deg * agg[t];  out[t] = mean[t] @ W
__global__ void __launch_bounds__(256)
k_hier_fwd_epilogue(const float* __restrict__ inv_deg, float* __restrict__ mean, const float* __restrict__ weight,
                    int n_source, int n_target, int f_in, int f_out, float* __restrict__ out) {
    extern __shared__ float sm[];  // weight [f_in*f_out] | rows [8][f_in]
    float* sw = sm;
    float* rows = sm + f_in * f_out;
    for (int i = threadIdx.x; i < f_in * f_out; i += blockDim.x) sw[i] = weight[i];
    const int w = warp_id(), lane = lane_id();
    const int t = blockIdx.x * 8 + w;
    if (t < n_target) {
        const float sc = inv_deg[n_source + t];
        for (int f = lane; f < f_in; f += 32) {
            float m = mean[int64_t(t) * f_in + f] * sc;
            mean[int64_t(t) * f_in + f] = m;
            rows[w * f_in + f] = m;
        }
    }
    __syncthreads();
    if (t < n_target) {
        for (int o = lane; o < f_out; o += 32) {
            float acc = 0.f;
            for (int f = 0; f < f_in; ++f) acc = fmaf(rows[w * f_in + f], sw[f * f_out + o], acc);
            out[int64_t(t) * f_out + o] = acc;
        }
    }
}

// hierarchy backward prologue: d_mean_full[n_source + t] = inv_deg[t] * (gout[t] @ W^T); d_mean[t] = gout[t] @ W^T
__global__ void __launch_bounds__(256)
k_hier_bwd_prologue(const float* __restrict__ inv_deg, const float* __restrict__ gout, const float* __restrict__ weight,
                    int n_source, int n_target, int f_in, int f_out, float* __restrict__ d_mean_full) {
    extern __shared__ float sm[];  // weight | gout rows [8][f_out]
    float* sw = sm;
    float* rows = sm + f_in * f_out;
    for (int i = threadIdx.x; i < f_in * f_out; i += blockDim.x) sw[i] = weight[i];
    const int w = warp_id(), lane = lane_id();
    const int t = blockIdx.x * 8 + w;
    if (t < n_target)
        for (int o = lane; o < f_out; o += 32) rows[w * f_out + o] = gout[int64_t(t) * f_out + o];
    __syncthreads();
    if (t < n_target) {
        const float sc = inv_deg[n_source + t];
        for (int f = lane; f < f_in; f += 32) {
            float acc = 0.f;
            for (int o = 0; o < f_out; ++o) acc = fmaf(rows[w * f_out + o], sw[f * f_out + o], acc);
            d_mean_full[int64_t(n_source + t) * f_in + f] = acc * sc;
        }
    }
}

int atb_launch(const float* A, const float* Bm, int K, int M, int N, float* out, float* partial_ws, cudaStream_t s);
size_t atb_ws_floats(int M, int N);

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_gcn_norm(const void* plan, int64_t n_entries, int64_t n_nodes, float* dis, void* stream) {
    TIPB_CHECK_ARG(plan && dis, "gcn_norm: NULL argument");
    CsrView v = csr_view(plan, n_entries, n_nodes, 1);
    k_gcn_norm<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(v.deg, (int)n_nodes, dis);
    TIPB_CHECK_LAUNCH("gcn_norm");
    return TIPB_OK;
}

int tipb_gcn_spmm(const void* plan, int64_t n_entries, int64_t n_nodes, const float* dis_out, const float* dis_in,
                  const float* x, const float* bias, int f, int relu, float* out, void* stream) {
    TIPB_CHECK_ARG(plan && x && out, "gcn_spmm: NULL argument");
    CsrView v = csr_view(plan, n_entries, n_nodes, 1);
    return node_aggregate_launch(v, dis_out, dis_in, x, bias, 0, (int)n_nodes, f, 1, relu, out, (cudaStream_t)stream);
}

size_t tipb_hier_workspace_bytes(int64_t n_source, int64_t n_target, int f_in, int f_out) {
    return (size_t(n_source + n_target) * f_in + atb_ws_floats(f_in, f_out)) * 4 + 1024;
}

int tipb_hier_fwd(const void* plan_by_dst, int64_t n_entries, int64_t n_source, int64_t n_target, const float* x,
                  const float* weight, int f_in, int f_out, float* mean, float* out, void* stream) {
    TIPB_CHECK_ARG(plan_by_dst && x && weight && mean && out, "hier_fwd: NULL argument");
    TIPB_CHECK_ARG(size_t(f_in) * f_out + 8 * f_in <= 11 * 1024, "hier_fwd: weight too large for the fused epilogue");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = n_source + n_target;
    CsrView v = csr_view(plan_by_dst, n_entries, n, 1);
    int rc = node_aggregate_launch(v, nullptr, nullptr, x, nullptr, (int)n_source, (int)n, f_in, 0, 0, mean, s);
    if (rc) return rc;
    const size_t smem = (size_t(f_in) * f_out + 8 * f_in) * sizeof(float);
    k_hier_fwd_epilogue<<<(unsigned)ceil_div(n_target, 8), 256, smem, s>>>(v.inv_deg, mean, weight, (int)n_source,
                                                                           (int)n_target, f_in, f_out, out);
    TIPB_CHECK_LAUNCH("hier_fwd");
    return TIPB_OK;
}

int tipb_hier_bwd(const void* plan_by_src, int64_t n_entries, int64_t n_source, int64_t n_target,
                  const float* inv_deg_dst, const float* mean, const float* weight, const float* grad_out, int f_in,
                  int f_out, float* d_x, float* d_weight, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(plan_by_src && inv_deg_dst && mean && weight && grad_out && d_x && d_weight && ws, "hier_bwd: NULL argument");
    TIPB_CHECK_ARG(size_t(f_in) * f_out + 8 * f_out <= 11 * 1024, "hier_bwd: weight too large for the fused prologue");
    TIPB_CHECK_ARG(ws_bytes >= tipb_hier_workspace_bytes(n_source, n_target, f_in, f_out), "hier_bwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = n_source + n_target;
    CsrView v = csr_view(plan_by_src, n_entries, n, 1);
    Carver c(ws);
    float* d_mean_full = c.take<float>(size_t(n) * f_in);
    float* partial = c.take<float>(atb_ws_floats(f_in, f_out));
    TIPB_CHECK_CUDA(cudaMemsetAsync(d_mean_full, 0, size_t(n_source) * f_in * sizeof(float), s));
    const size_t smem = (size_t(f_in) * f_out + 8 * f_out) * sizeof(float);
    k_hier_bwd_prologue<<<(unsigned)ceil_div(n_target, 8), 256, smem, s>>>(inv_deg_dst, grad_out, weight, (int)n_source,
                                                                           (int)n_target, f_in, f_out, d_mean_full);
    // d_x[s] = sum_{t: s->t} inv_deg[t] * d_mean[t]
    int rc = node_aggregate_launch(v, nullptr, nullptr, d_mean_full, nullptr, 0, (int)n, f_in, 0, 0, d_x, s);
    if (rc) return rc;
    // d_weight = mean^T @ gout
    if (f_out == 4 || f_out == 8 || f_out == 16 || f_out == 32 || f_out == 64 || f_out == 128)
        return atb_launch(mean, grad_out, (int)n_target, f_in, f_out, d_weight, partial, s);
    set_last_error("hier_bwd: f_out=%d must be a power of two in [4,128]", f_out);
    return TIPB_ERR_UNSUPPORTED;
}
}
