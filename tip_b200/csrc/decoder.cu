// DistMult decoder (reference src/layers.py:581-595), the TIP loss (src/layers.py:335-340) and
// their gradients (north_star item 4: fused gather-bilinear-sigmoid-BCE).
//
//   value_e = sum_k z[i_e,k] z[j_e,k] w[r_e,k]        score_e = sigmoid(value_e)
//   loss    = -mean(log(score_pos + 1e-13)) - mean(log(1 - score_neg + 1e-13))
//
// The gradient with respect to z is a scatter over BOTH endpoints of every pair.  To keep it free
// of atomics, every edge set is indexed by a *doubled* typed CSR: each pair (i,j,r) is listed under
// (node i, relation r) with other=j and again under (node j, relation r) with other=i.  One warp
// owns one (node, relation) segment: u = z[node] * w[rel] is loop-invariant, each entry costs one
// 64-byte gather of z[other] (from shared memory), one dot product, and -- for the fused loss -- the
// sigmoid/log/derivative evaluated ONCE per entry with all 32 lanes busy (entries are processed 32 at
// a time; the per-entry scalars are transposed across the warp with shuffles).  The segment leaves
//   acc[s,:] = sum_e g_e z[other_e,:]
// from which  d_z[n] = sum_{s in node n} w[rel_s] * acc[s]   (node-major reduction) and
//             d_w[r] = 1/2 sum_{s: rel_s = r} z[node_s] * acc[s]   (relation-major reduction; every pair
//                                                                 was visited from both ends).
#include <stdlib.h>

#include "common.cuh"
#include "reduce.cuh"

namespace tipb {

constexpr float TIP_EPS = 1e-13f;  // reference src/layers.py:15

enum { DEC_MODE_POS = 0, DEC_MODE_NEG = 1, DEC_MODE_GRAD = 2 };

__device__ __forceinline__ float sigmoidf_ref(float v) { return 1.0f / (1.0f + expf(-v)); }
// SFU versions for the fused loss kernel (MUFU.EX2 / MUFU.RCP / MUFU.LG2): a few ulp, far inside rtol 1e-4
__device__ __forceinline__ float sigmoidf_fast(float v) { return __frcp_rn(1.0f + __expf(-v)); }

// ------------------------------------------------------------------------------------------------
// plain forward in the caller's edge order (the module's public forward)
template <int LPR>
__global__ void __launch_bounds__(256)
k_decoder_fwd(const float4* __restrict__ z, const float4* __restrict__ w, const int64_t* __restrict__ edge_index,
              const int64_t* __restrict__ edge_type, int64_t n_edges, int n_nodes, int n_rel, int apply_sigmoid,
              float* __restrict__ out) {
    constexpr int G = 32 / LPR;
    const int lane = lane_id(), g = lane / LPR, l = lane % LPR;
    const int64_t n_groups = (int64_t(gridDim.x) * blockDim.x) / LPR;
    for (int64_t e = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / LPR; e < ((n_edges + G - 1) / G) * G;
         e += n_groups) {
        const bool ok = e < n_edges;
        float p = 0.f;
        if (ok) {
            int64_t i = edge_index[e], j = edge_index[n_edges + e], r = edge_type[e];
            bool in = i >= 0 && i < n_nodes && j >= 0 && j < n_nodes && r >= 0 && r < n_rel;
            if (in) {
                float4 a = z[i * LPR + l], b = z[j * LPR + l], c = w[r * LPR + l];
                p = (a.x * b.x) * c.x + (a.y * b.y) * c.y + (a.z * b.z) * c.z + (a.w * b.w) * c.w;
            } else {
                p = __int_as_float(0x7fc00000);  // NaN marks an out-of-range index
            }
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
        if (ok && l == 0) out[e] = apply_sigmoid ? sigmoidf_ref(p) : p;
    }
    (void)g;
}

// ------------------------------------------------------------------------------------------------
// segment kernel.  smem: z [n_nodes * LPR] float4 when it fits (STAGED), else rows come from L2 through the
// read-only path (scaled graphs: 10^4 drugs x 16 floats = 640 KB)
template <int LPR, int MODE, bool STAGED>
__global__ void __launch_bounds__(512)
k_decoder_seg(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node, const int* __restrict__ seg_rel,
              const int* __restrict__ other, const int* __restrict__ eid, const int* __restrict__ counts,
              const float4* __restrict__ z, const float4* __restrict__ w, const float* __restrict__ grad_out,
              int n_nodes, int n_edges, int apply_sigmoid, float inv_count, float4* __restrict__ acc_seg,
              float4* __restrict__ zacc_seg, float* __restrict__ loss_part) {
    extern __shared__ float4 s_z[];
    constexpr int G = 32 / LPR;
    const int lane = lane_id(), g = lane / LPR, l = lane % LPR;
    if (STAGED) {
        for (int i = threadIdx.x; i < n_nodes * LPR; i += blockDim.x) s_z[i] = z[i];
        __syncthreads();
    }
    const float4* __restrict__ zsrc = STAGED ? s_z : z;

    const int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float loss = 0.f;

    for (int s = warp_global; s < S; s += n_warps) {
        const int beg = seg_ptr[s], end = seg_ptr[s + 1];
        const int node = seg_node[s], rel = seg_rel[s];
        const float4 zn = zsrc[node * LPR + l];
        const float4 wr = w[rel * LPR + l];
        const float4 u = f4_mul(zn, wr);
        float4 acc = f4_zero();

        for (int base = beg; base < end; base += 32) {
            const int cnt = min(32, end - base);
            const int my = base + lane;
            const int idx = my < end ? ld_stream_i32(other + my) : 0;
            float go = 0.f;
            if (MODE == DEC_MODE_GRAD && my < end) {
                int id = eid[my];
                go = grad_out[id >= n_edges ? id - n_edges : id];
            }
            // phase A: one dot product per entry, entry k handled by group k % G in round k / G
            float4 zj[LPR];     // LPR == number of rounds: 32 / G
            float vmine = 0.f;  // value of entry `lane`
#pragma unroll
            for (int r = 0; r < LPR; ++r) {
                const int k = r * G + g;
                const int j = __shfl_sync(FULL, idx, k);
                zj[r] = zsrc[j * LPR + l];
                float p = f4_dot(u, zj[r]);
#pragma unroll
                for (int o = LPR >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
                // entry k's value sits on every lane of group g; lane k fetches it from lane (k % G) * LPR
                const float pv = __shfl_sync(FULL, p, (lane % G) * LPR);
                if (lane / G == r) vmine = pv;
            }
            // phase B: all 32 lanes evaluate the scalar chain of their own entry
            float gmine = 0.f;
            if (lane < cnt) {
                if (MODE == DEC_MODE_GRAD) {
                    if (apply_sigmoid) {
                        const float sg = sigmoidf_fast(vmine);
                        gmine = go * sg * (1.f - sg);
                    } else {
                        gmine = go;
                    }
                } else {
                    const float sg = sigmoidf_fast(vmine);
                    if (MODE == DEC_MODE_POS) {
                        loss -= __logf(sg + TIP_EPS);
                        gmine = -__fdividef(sg * (1.f - sg), sg + TIP_EPS) * inv_count;
                    } else {
                        const float om = 1.f - sg;
                        loss -= __logf(om + TIP_EPS);
                        gmine = __fdividef(sg * om, om + TIP_EPS) * inv_count;
                    }
                }
            }
            // phase C: acc += g_e * z[other_e]
#pragma unroll
            for (int r = 0; r < LPR; ++r) {
                const float ge = __shfl_sync(FULL, gmine, r * G + g);
                acc = f4_fma(ge, zj[r], acc);
            }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(FULL, acc.x, o);
            acc.y += __shfl_xor_sync(FULL, acc.y, o);
            acc.z += __shfl_xor_sync(FULL, acc.z, o);
            acc.w += __shfl_xor_sync(FULL, acc.w, o);
        }
        if (g == 0) {
            acc_seg[int64_t(s) * LPR + l] = f4_mul(acc, wr);
            zacc_seg[int64_t(s) * LPR + l] = f4_mul(acc, zn);
        }
    }
    if (MODE != DEC_MODE_GRAD) {
        loss = warp_sum(loss);
        if (lane == 0) loss_part[warp_global] = loss;
    }
}

// d_z[n, :] (+)= sum_{s in node n} wacc_seg[s, :]
// One CTA per node; NODE_REDUCE_THREADS / dim row groups, four rows in flight per group, fixed summation order.
constexpr int NODE_REDUCE_THREADS = 512;
__global__ void __launch_bounds__(NODE_REDUCE_THREADS)
k_decoder_node_reduce(const int* __restrict__ node_ptr, const int* __restrict__ listing, const int* __restrict__ counts,
                      const float* __restrict__ wacc_seg, int dim, int accumulate, float* __restrict__ d_z) {
    __shared__ float part[NODE_REDUCE_THREADS];
    const int n = blockIdx.x;
    const int sb = node_ptr[n], se = node_ptr[n + 1];
    const bool listed = counts[TIPB_CSR_COUNT_REL_MAJOR] != 0;  // relation-major plan: node_ptr indexes the listing
    const int per = NODE_REDUCE_THREADS / dim;  // dim <= 128 and a power of two
    const int k = threadIdx.x % dim, sg = threadIdx.x / dim;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = sb + sg;
    for (; i + 3 * per < se; i += 4 * per) {
        const int s0 = listed ? listing[i] : i, s1 = listed ? listing[i + per] : i + per;
        const int s2 = listed ? listing[i + 2 * per] : i + 2 * per, s3 = listed ? listing[i + 3 * per] : i + 3 * per;
        a0 += wacc_seg[int64_t(s0) * dim + k];
        a1 += wacc_seg[int64_t(s1) * dim + k];
        a2 += wacc_seg[int64_t(s2) * dim + k];
        a3 += wacc_seg[int64_t(s3) * dim + k];
    }
    for (; i < se; i += per) a0 += wacc_seg[int64_t(listed ? listing[i] : i) * dim + k];
    part[threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (threadIdx.x < dim) {
        float t = 0.f;
        for (int q = 0; q < per; ++q) t += part[q * dim + threadIdx.x];
        float* dst = d_z + int64_t(n) * dim + threadIdx.x;
        *dst = accumulate ? *dst + t : t;
    }
}

__global__ void __launch_bounds__(1024)
k_loss_reduce(const float* __restrict__ loss_part, int n, float scale, int accumulate, float* __restrict__ loss_out) {
    __shared__ float sw[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) a += loss_part[i];
    a = warp_sum(a);
    if (lane_id() == 0) sw[warp_id()] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += sw[w];
        t *= scale;
        loss_out[0] = accumulate ? loss_out[0] + t : t;
    }
}

__global__ void k_add_inplace(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// full sweep: out[r, i, j] = act( sum_k z[i,k] w[r,k] z[j,k] )     (BASELINE.json config 5)
__global__ void __launch_bounds__(256)
k_decoder_sweep(const float* __restrict__ z, const float* __restrict__ w, int n_nodes, int dim, int apply_sigmoid,
                float* __restrict__ out) {
    extern __shared__ float sm[];  // u [8][dim] for this CTA's 8 rows i
    const int r = blockIdx.y;
    const int i0 = blockIdx.x * 8;
    for (int t = threadIdx.x; t < 8 * dim; t += blockDim.x) {
        int ii = i0 + t / dim, k = t % dim;
        sm[t] = ii < n_nodes ? z[int64_t(ii) * dim + k] * w[int64_t(r) * dim + k] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_nodes; j += blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < dim; ++k) {
            const float zj = z[int64_t(j) * dim + k];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(sm[q * dim + k], zj, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int ii = i0 + q;
            if (ii < n_nodes) {
                float v = acc[q];
                if (apply_sigmoid) v = 1.0f / (1.0f + __expf(-v));
                out[(int64_t(r) * n_nodes + ii) * n_nodes + j] = v;
            }
        }
    }
}

// the layout src/utils.py:17-23 (to_bidirection) produces: every relation range is [pairs..., the same pairs with
// rows swapped...]; flag stays 1 iff that holds for every relation (and the ranges tile [0, n_edges))
__global__ void k_mirror_init(int* flag) { *flag = 1; }
__global__ void __launch_bounds__(256)
k_mirror_check(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ range_list, int64_t n_edges,
               int* __restrict__ flag) {
    const int r = blockIdx.x;
    const int64_t beg = range_list[2 * r], end = range_list[2 * r + 1];
    const int64_t prev_end = r == 0 ? 0 : range_list[2 * r - 1];
    bool ok = beg == prev_end && end >= beg && end <= n_edges && ((end - beg) & 1) == 0;
    if (r == int(gridDim.x) - 1) ok = ok && end == n_edges;
    if (ok) {
        const int64_t half = (end - beg) >> 1;
        for (int64_t e = beg + threadIdx.x; e < beg + half; e += 256)
            ok = ok && edge_index[e] == edge_index[n_edges + e + half] && edge_index[n_edges + e] == edge_index[e + half];
    }
    if (!ok) atomicAnd(flag, 0);
}

// Register-tiled sweep for the embedding widths the model uses (DIM in {4, 8, 16, 32}) when z fits in shared memory.
// The sweep is 11.5 GFLOP of fp32 FMA against 1.43 GB of output at config 5, so it has to run the FMA pipe and the
// DRAM write stream side by side: a CTA takes 64 rows i of one relation, stages u_i = z_i * w_r and the whole z
// (row pitch DIM + 4 floats: conflict-free LDS.128 for consecutive j); a warp owns 8 rows, a lane 4 columns
// j = jb + lane + 32 c, i.e. an 8 x 4 accumulator tile per thread fed by 12 LDS.128 per 128 FMA; every store
// instruction writes 32 consecutive j of one row.
constexpr int SWEEP_ROWS = 64;
template <int DIM>
__global__ void __launch_bounds__(256)
k_decoder_sweep_tiled(const float* __restrict__ z, const float* __restrict__ w, int n_nodes, int apply_sigmoid,
                      float* __restrict__ out) {
    constexpr int Q = DIM / 4, PITCH4 = Q + 1;  // float4 per row, staged pitch in float4
    extern __shared__ float4 sw4[];
    float4* zs = sw4;                       // [n_nodes][PITCH4]
    float4* us = sw4 + size_t(n_nodes) * PITCH4;  // [SWEEP_ROWS][Q]
    const int r = blockIdx.y, i0 = blockIdx.x * SWEEP_ROWS;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    const float4* w4 = reinterpret_cast<const float4*>(w) + size_t(r) * Q;
    for (int t = threadIdx.x; t < n_nodes * Q; t += 256) zs[(t / Q) * PITCH4 + (t % Q)] = z4[t];
    for (int t = threadIdx.x; t < SWEEP_ROWS * Q; t += 256) {
        const int i = i0 + t / Q;
        us[t] = i < n_nodes ? f4_mul(z4[size_t(i) * Q + (t % Q)], w4[t % Q]) : f4_zero();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* urow = us + warp * 8 * Q;
    for (int jb = 0; jb < n_nodes; jb += 128) {
        float acc[8][4];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;
        int jc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) jc[c] = min(jb + lane + 32 * c, n_nodes - 1);
#pragma unroll 1
        for (int k4 = 0; k4 < Q; ++k4) {
            float4 zc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) zc[c] = zs[jc[c] * PITCH4 + k4];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 u = urow[q * Q + k4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[q][c] = fmaf(u.x, zc[c].x, acc[q][c]);
                    acc[q][c] = fmaf(u.y, zc[c].y, acc[q][c]);
                    acc[q][c] = fmaf(u.z, zc[c].z, acc[q][c]);
                    acc[q][c] = fmaf(u.w, zc[c].w, acc[q][c]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + warp * 8 + q;
            if (i < n_nodes) {
                float* row = out + (size_t(r) * n_nodes + i) * n_nodes;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = jb + lane + 32 * c;
                    float v = acc[q][c];
                    if (apply_sigmoid) v = 1.0f / (1.0f + __expf(-v));
                    if (j < n_nodes) row[j] = v;
                }
            }
        }
    }
}

template <int DIM>
static int sweep_tiled_launch(const float* z, const float* w, int64_t n_nodes, int64_t n_rel, int apply_sigmoid, float* out,
                              cudaStream_t s) {
    const size_t smem = (size_t(n_nodes) * (DIM / 4 + 1) + size_t(SWEEP_ROWS) * (DIM / 4)) * sizeof(float4);
    auto kern = k_decoder_sweep_tiled<DIM>;
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    dim3 grid((unsigned)ceil_div(n_nodes, SWEEP_ROWS), (unsigned)n_rel);
    kern<<<grid, 256, smem, s>>>(z, w, (int)n_nodes, apply_sigmoid, out);
    return TIPB_OK;
}

static int dec_seg_grid() { return sm_count() * 2; }

struct DecWs { float *acc_seg, *zacc_seg, *loss_part, *dw_tmp; };
static size_t dec_ws_bytes(int64_t seg_cap, int64_t n_rel, int dim) {
    return (2 * size_t(seg_cap) * dim + size_t(dec_seg_grid()) * 16 + size_t(n_rel) * dim) * 4 + 2048;
}

template <int LPR>
static int decoder_seg_run(const CsrView& v, int mode, const float* z, const float* w, const float* grad_out,
                           int64_t n_edges, int apply_sigmoid, int accumulate, int mult, float* loss_out, float* d_z,
                           float* d_w, void* ws, cudaStream_t s) {
    const int dim = LPR * 4;
    Carver c(ws);
    float* acc_seg = c.take<float>(size_t(v.seg_cap) * dim);
    float* zacc_seg = c.take<float>(size_t(v.seg_cap) * dim);
    const int grid = dec_seg_grid();
    const int n_warps = grid * 16;
    float* loss_part = c.take<float>(n_warps);
    float* dw_tmp = c.take<float>(size_t(v.n_rel) * dim);
    const bool staged = size_t(v.n_nodes) * dim * sizeof(float) + 1024 <= size_t(max_smem_optin());
    const size_t smem = staged ? size_t(v.n_nodes) * dim * sizeof(float) : 0;
    // mult = 1: doubled plan (every pair listed under both endpoints); mult = 2: one listing per directed edge of a
    // mirrored edge set (each (node, relation) segment of the doubled plan would hold every neighbour twice)
    const float inv_count = n_edges > 0 ? float(mult) / float(n_edges) : 0.f;
    int rc;
#define RUN1(MODEV, STG)                                                                                            \
    {                                                                                                               \
        auto kern = k_decoder_seg<LPR, MODEV, STG>;                                                                 \
        if ((rc = ensure_dyn_smem((const void*)kern, smem))) return rc;                                             \
        kern<<<grid, 512, smem, s>>>(v.seg_ptr, v.seg_node, v.seg_rel, v.other, v.eid, v.counts, (const float4*)z,  \
                                     (const float4*)w, grad_out, (int)v.n_nodes, (int)n_edges, apply_sigmoid,       \
                                     inv_count, (float4*)acc_seg, (float4*)zacc_seg, loss_part);                    \
    }
#define RUN(MODEV) { if (staged) RUN1(MODEV, true) else RUN1(MODEV, false) }
    if (mode == DEC_MODE_POS) RUN(DEC_MODE_POS)
    else if (mode == DEC_MODE_NEG) RUN(DEC_MODE_NEG)
    else RUN(DEC_MODE_GRAD)
#undef RUN
#undef RUN1
    k_decoder_node_reduce<<<(unsigned)v.n_nodes, NODE_REDUCE_THREADS, 0, s>>>(v.node_ptr, v.rel_seg, v.counts, acc_seg, dim, accumulate, d_z);
    if (accumulate) {
        k_rel_reduce<<<(unsigned)v.n_rel, REL_REDUCE_THREADS, 0, s>>>(v.rel_seg_ptr, v.rel_seg, v.counts, zacc_seg, dim, 0.5f, dw_tmp);
        k_add_inplace<<<(unsigned)ceil_div(v.n_rel * dim, 256), 256, 0, s>>>(d_w, dw_tmp, v.n_rel * dim);
    } else {
        k_rel_reduce<<<(unsigned)v.n_rel, REL_REDUCE_THREADS, 0, s>>>(v.rel_seg_ptr, v.rel_seg, v.counts, zacc_seg, dim, 0.5f, d_w);
    }
    if (mode != DEC_MODE_GRAD)
        k_loss_reduce<<<1, 1024, 0, s>>>(loss_part, n_warps, 0.5f * inv_count, accumulate, loss_out);  // inv_count has mult
    TIPB_CHECK_LAUNCH("decoder_seg");
    return TIPB_OK;
}

static int decoder_seg_dispatch(const void* plan, int mode, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                                const float* z, const float* w, const float* grad_out, int dim, int apply_sigmoid,
                                int accumulate, int mult, float* loss_out, float* d_z, float* d_w, void* ws,
                                size_t ws_bytes, cudaStream_t s) {
    CsrView v = csr_view(plan, mult == 2 ? n_edges : 2 * n_edges, n_nodes, n_rel);
    TIPB_CHECK_ARG(ws_bytes >= dec_ws_bytes(v.seg_cap, n_rel, dim), "decoder: workspace too small");
    switch (dim) {
        case 4: return decoder_seg_run<1>(v, mode, z, w, grad_out, n_edges, apply_sigmoid, accumulate, mult, loss_out, d_z, d_w, ws, s);
        case 8: return decoder_seg_run<2>(v, mode, z, w, grad_out, n_edges, apply_sigmoid, accumulate, mult, loss_out, d_z, d_w, ws, s);
        case 16: return decoder_seg_run<4>(v, mode, z, w, grad_out, n_edges, apply_sigmoid, accumulate, mult, loss_out, d_z, d_w, ws, s);
        case 32: return decoder_seg_run<8>(v, mode, z, w, grad_out, n_edges, apply_sigmoid, accumulate, mult, loss_out, d_z, d_w, ws, s);
    }
    set_last_error("decoder: dim=%d not in {4,8,16,32} (the Python layer pads)", dim);
    return TIPB_ERR_UNSUPPORTED;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_decoder_fwd(const float* z, const float* weight, const int64_t* edge_index, const int64_t* edge_type,
                     int64_t n_edges, int64_t n_nodes, int64_t n_rel, int dim, int apply_sigmoid, float* out,
                     void* stream) {
    TIPB_CHECK_ARG(z && weight && out && (n_edges == 0 || (edge_index && edge_type)), "decoder_fwd: NULL argument");
    if (n_edges == 0) return TIPB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = sm_count() * 8;
#define FWD(LPRV)                                                                                                  \
    k_decoder_fwd<LPRV><<<blocks, 256, 0, s>>>((const float4*)z, (const float4*)weight, edge_index, edge_type, n_edges, \
                                               (int)n_nodes, (int)n_rel, apply_sigmoid, out)
    switch (dim) {
        case 4: FWD(1); break;
        case 8: FWD(2); break;
        case 16: FWD(4); break;
        case 32: FWD(8); break;
        case 64: FWD(16); break;
        case 128: FWD(32); break;
        default:
            set_last_error("decoder_fwd: dim=%d not in {4,...,128}", dim);
            return TIPB_ERR_UNSUPPORTED;
    }
#undef FWD
    TIPB_CHECK_LAUNCH("decoder_fwd");
    return TIPB_OK;
}

size_t tipb_decoder_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t n_rel, int dim) {
    int64_t cap = n_nodes * n_rel, ent = 2 * n_edges;
    int64_t seg_cap = ent < cap ? ent : cap;
    if (seg_cap < 1) seg_cap = 1;
    return dec_ws_bytes(seg_cap, n_rel, dim);
}

int tipb_decoder_bwd(const void* plan_doubled, int64_t n_edges, int64_t n_nodes, int64_t n_rel, const float* z,
                     const float* weight, const float* grad_out, int dim, int apply_sigmoid, float* d_z,
                     float* d_weight, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(plan_doubled && z && weight && grad_out && d_z && d_weight && ws, "decoder_bwd: NULL argument");
    return decoder_seg_dispatch(plan_doubled, DEC_MODE_GRAD, n_edges, n_nodes, n_rel, z, weight, grad_out, dim,
                                apply_sigmoid, 0, 1, nullptr, d_z, d_weight, ws, ws_bytes, (cudaStream_t)stream);
}

int tipb_decoder_bce_fused(const void* plan_doubled, int64_t n_edges, int64_t n_nodes, int64_t n_rel, const float* z,
                           const float* weight, int dim, int sign, int accumulate, float* loss_out, float* d_z,
                           float* d_weight, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(plan_doubled && z && weight && loss_out && d_z && d_weight && ws, "decoder_bce_fused: NULL argument");
    TIPB_CHECK_ARG(sign == 1 || sign == -1, "decoder_bce_fused: sign must be +1 (positives) or -1 (negatives)");
    return decoder_seg_dispatch(plan_doubled, sign > 0 ? DEC_MODE_POS : DEC_MODE_NEG, n_edges, n_nodes, n_rel, z, weight,
                                nullptr, dim, 1, accumulate, 1, loss_out, d_z, d_weight, ws, ws_bytes, (cudaStream_t)stream);
}

int tipb_decoder_bce_fused_mirrored(const void* plan_by_target, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                                    const float* z, const float* weight, int dim, int sign, int accumulate,
                                    float* loss_out, float* d_z, float* d_weight, void* ws, size_t ws_bytes,
                                    void* stream) {
    TIPB_CHECK_ARG(plan_by_target && z && weight && loss_out && d_z && d_weight && ws,
                   "decoder_bce_fused_mirrored: NULL argument");
    TIPB_CHECK_ARG(sign == 1 || sign == -1, "decoder_bce_fused_mirrored: sign must be +1 or -1");
    return decoder_seg_dispatch(plan_by_target, sign > 0 ? DEC_MODE_POS : DEC_MODE_NEG, n_edges, n_nodes, n_rel, z, weight,
                                nullptr, dim, 1, accumulate, 2, loss_out, d_z, d_weight, ws, ws_bytes,
                                (cudaStream_t)stream);
}

int tipb_edges_mirrored(const int64_t* edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_rel,
                        int32_t* flag_out, void* stream) {
    TIPB_CHECK_ARG(edge_index && range_list && flag_out, "edges_mirrored: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    k_mirror_init<<<1, 1, 0, s>>>(flag_out);
    if (n_rel > 0)
        k_mirror_check<<<(unsigned)n_rel, 256, 0, s>>>(edge_index, range_list, n_edges, flag_out);
    TIPB_CHECK_LAUNCH("edges_mirrored");
    return TIPB_OK;
}

int tipb_decoder_sweep_status(void) {
    int flag = 0;
    if (cudaMemcpy(&flag, sweep_tc_error_flag(), sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (flag) cudaMemset(sweep_tc_error_flag(), 0, sizeof(int));
    return flag;
}

int tipb_decoder_sweep(const float* z, const float* weight, int64_t n_nodes, int64_t n_rel, int dim, int apply_sigmoid,
                       float* out, void* stream) {
    TIPB_CHECK_ARG(z && weight && out, "decoder_sweep: NULL argument");
    TIPB_CHECK_ARG(dim >= 1 && dim <= 1024 && n_rel <= 65535, "decoder_sweep: dim/n_rel out of range");
    // tcgen05 path (sweep_tc.cu): the contraction as one bf16x3-split GEMM with TMEM accumulators, write-bound
    static const bool tc_off = getenv("TIPB_SWEEP_TC") && getenv("TIPB_SWEEP_TC")[0] == '0';
    if (!tc_off && sweep_tc_supported(n_nodes, n_rel, dim) && (reinterpret_cast<uintptr_t>(z) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(weight) & 15) == 0)
        return sweep_tc_run(z, weight, n_nodes, n_rel, dim, apply_sigmoid, out, (cudaStream_t)stream);
    const bool tiled_ok = (dim == 4 || dim == 8 || dim == 16 || dim == 32) && n_nodes < (int64_t(1) << 24) &&
                          (size_t(n_nodes) * (dim / 4 + 1) + size_t(SWEEP_ROWS) * (dim / 4)) * sizeof(float4) + 1024 <=
                              size_t(max_smem_optin()) &&
                          (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0;
    if (tiled_ok) {
        int rc = TIPB_OK;
        switch (dim) {
            case 4: rc = sweep_tiled_launch<4>(z, weight, n_nodes, n_rel, apply_sigmoid, out, (cudaStream_t)stream); break;
            case 8: rc = sweep_tiled_launch<8>(z, weight, n_nodes, n_rel, apply_sigmoid, out, (cudaStream_t)stream); break;
            case 16: rc = sweep_tiled_launch<16>(z, weight, n_nodes, n_rel, apply_sigmoid, out, (cudaStream_t)stream); break;
            default: rc = sweep_tiled_launch<32>(z, weight, n_nodes, n_rel, apply_sigmoid, out, (cudaStream_t)stream); break;
        }
        if (rc) return rc;
    } else {
        dim3 grid((unsigned)ceil_div(n_nodes, 8), (unsigned)n_rel);
        k_decoder_sweep<<<grid, 256, 8 * dim * sizeof(float), (cudaStream_t)stream>>>(z, weight, (int)n_nodes, dim,
                                                                                      apply_sigmoid, out);
    }
    TIPB_CHECK_LAUNCH("decoder_sweep");
    return TIPB_OK;
}
}
