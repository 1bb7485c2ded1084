// Small dense products and the operators of the reference's ablation models, plus the
// dense pieces of the P-P encoder that round 1 left to cuBLAS / eager torch (GCNConv.lin of the second layer, its
// backward, the ReLU mask and the bias gradient; reference src/layers.py:391-395).
//
//   tipb_gemm            C = A B or A B^T (+ bias, + ReLU), fp32, row-major, small shapes (<= 2*10^4 x 128)
//   tipb_gemm_tn         C = A^T B, split-K in a fixed order (weight gradients)
//   tipb_relu_grad_colsum  g = gy masked by (relu_out > 0) and the column sums of g (bias gradient) in one call
//   tipb_transpose       out = in^T (GCNConv.lin of an identity feature matrix: x' = lin.weight^T)
//   tipb_drug_input_fwd/_bwd   FMEncoder's glue: embed / d_norm, then cat | add with the hierarchy output
//   tipb_scale2          two buffers times one device scalar (the loss gradient arriving at the fused loss kernels)
//   tipb_nn_decoder_*    NNDecoder (src/layers.py:598-637): sigmoid( relu(z_i W1) . w1_r + relu(z_j W2) . w2_r );
//                        the per-node hidden layers and the [N x R] tables A = H1 w1^T, B = H2 w2^T are dense products,
//                        the per-edge work is a gather of two scalars
//   tipb_spmm_values     out[i] = sum_k val_ik dense[k]  over a typed CSR: general sparse drug features
//                        (data/utils.py:117-132: identity + mono side-effect columns) as the front-end of `embed`
// CUDA cores only: every product here is far below 1 GFLOP.
#include "common.cuh"

namespace tipb {

int atb_launch(const float* A, const float* Bm, int K, int M, int N, float* out, float* partial_ws, cudaStream_t s,
               int slice_cap);
size_t atb_ws_floats(int M, int N);

constexpr int GT = 64, GK = 16;     // CTA tile 64 x 64, K chunk 16, 256 threads, 4 x 4 outputs per thread

// C[M,N] = A[M,K] * (TRANS_B ? B[N,K]^T : B[K,N])  (+ bias[N]) (+ ReLU)
template <bool TRANS_B>
__global__ void __launch_bounds__(256)
k_gemm(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias, int M, int N, int K,
       int relu, float* __restrict__ C) {
    __shared__ float As[GK][GT + 4];
    __shared__ float Bs[GK][GT + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GK) {
        for (int i = tid; i < GT * GK; i += 256) {
            const int r = i / GK, kk = i % GK;          // A: rows m0 + r, k-contiguous loads
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[size_t(m0 + r) * K + k0 + kk] : 0.f;
        }
        if (TRANS_B) {
            for (int i = tid; i < GT * GK; i += 256) {
                const int c = i / GK, kk = i % GK;      // B[N,K]: rows n0 + c
                Bs[kk][c] = (n0 + c < N && k0 + kk < K) ? B[size_t(n0 + c) * K + k0 + kk] : 0.f;
            }
        } else {
            for (int i = tid; i < GT * GK; i += 256) {
                const int kk = i / GT, c = i % GT;      // B[K,N]: n-contiguous loads
                Bs[kk][c] = (n0 + c < N && k0 + kk < K) ? B[size_t(k0 + kk) * N + n0 + c] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[size_t(m) * N + n] = v;
        }
    }
}

int gemm_launch(const float* A, const float* B, const float* bias, int64_t M, int64_t N, int64_t K, int trans_b, int relu,
                float* C, cudaStream_t s) {
    if (M == 0 || N == 0) return TIPB_OK;
    const dim3 grid((unsigned)ceil_div(N, GT), (unsigned)ceil_div(M, GT));
    if (trans_b) k_gemm<true><<<grid, 256, 0, s>>>(A, B, bias, (int)M, (int)N, (int)K, relu, C);
    else k_gemm<false><<<grid, 256, 0, s>>>(A, B, bias, (int)M, (int)N, (int)K, relu, C);
    TIPB_CHECK_LAUNCH("gemm");
    return TIPB_OK;
}

// g = gy masked by (mask_ref > 0) (written when g != NULL) and per-slice column sums for the bias gradient in one pass:
// CTA (x = 32 columns, y = row slice) walks its rows 8 at a time; partial[slice][n] are summed in slice order afterwards
constexpr int COLSUM_MAX_SLICES = 512;
__global__ void __launch_bounds__(256)
k_mask_colsum(const float* __restrict__ gy, const float* __restrict__ mask_ref, int64_t M, int N, int rows_per_slice,
              float* __restrict__ g, float* __restrict__ partial) {
    __shared__ float part[8][32];
    const int tx = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int64_t m0 = int64_t(blockIdx.y) * rows_per_slice, m1 = min(M, m0 + rows_per_slice);
    float a = 0.f;
    if (col < N)
        for (int64_t m = m0 + rg; m < m1; m += 8) {
            float v = gy[m * N + col];
            if (mask_ref && !(mask_ref[m * N + col] > 0.f)) v = 0.f;
            if (g) g[m * N + col] = v;
            a += v;
        }
    part[rg][tx] = a;
    __syncthreads();
    if (rg == 0 && col < N && partial) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][tx];
        partial[size_t(blockIdx.y) * N + col] = t;
    }
}
// one warp per column: lane l adds slices l, l + 32, ... in order, then a fixed shuffle tree
__global__ void __launch_bounds__(256)
k_colsum_finish(const float* __restrict__ partial, int n_slices, int N, float* __restrict__ db) {
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (col >= N) return;
    float t = 0.f;
    for (int k = lane; k < n_slices; k += 32) t += partial[size_t(k) * N + col];
    t = warp_sum(t);
    if (lane == 0) db[col] = t;
}

// ------------------------------------------------------------------------------------------------ NN decoder
// score[e] = act( A[i_e, r_e] + B[j_e, r_e] )    A, B = [n_nodes, n_rel] tables
__global__ void __launch_bounds__(256)
k_nn_decoder_fwd(const float* __restrict__ A, const float* __restrict__ B, const int64_t* __restrict__ edge_index,
                 const int64_t* __restrict__ edge_type, int64_t n_edges, int n_nodes, int n_rel, int apply_sigmoid,
                 float* __restrict__ out) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t i = edge_index[e], j = edge_index[n_edges + e], r = edge_type[e];
    if (i < 0 || i >= n_nodes || j < 0 || j >= n_nodes || r < 0 || r >= n_rel) {
        out[e] = __int_as_float(0x7fc00000);      // NaN marks an out-of-range index
        return;
    }
    const float v = A[i * n_rel + r] + B[j * n_rel + r];
    out[e] = apply_sigmoid ? 1.0f / (1.0f + expf(-v)) : v;
}
// per-edge gradient wrt the value: g[e] = gout[e] * s (1 - s)
__global__ void k_nn_decoder_gval(const float* __restrict__ gout, const float* __restrict__ score, int64_t n,
                                  int apply_sigmoid, float* __restrict__ g) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float s = score[e];
    g[e] = apply_sigmoid ? gout[e] * s * (1.f - s) : gout[e];
}
// dense[node, rel] = sum of g over the entries of segment (node, rel) of a typed CSR (zero elsewhere): one warp per
// segment, fixed order -- the scatter-add of the table gradients without float atomics
__global__ void __launch_bounds__(256)
k_seg_scalar_sum(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node, const int* __restrict__ seg_rel,
                 const int* __restrict__ eid, const int* __restrict__ counts, const float* __restrict__ g, int n_rel,
                 float* __restrict__ dense) {
    const int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    const int lane = lane_id();
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < S; s += (gridDim.x * blockDim.x) >> 5) {
        float a = 0.f;
        for (int p = seg_ptr[s] + lane; p < seg_ptr[s + 1]; p += 32) a += g[eid[p]];
        a = warp_sum(a);
        if (lane == 0) dense[int64_t(seg_node[s]) * n_rel + seg_rel[s]] = a;
    }
}

// ------------------------------------------------------------------------------------------------ valued SpMM
// out[n, :] = sum_{p in node n} val[eid[p]] * dense[other[p], :]       (F floats per row, F % 4 == 0; warp per row)
__global__ void __launch_bounds__(256)
k_spmm_values(const int* __restrict__ node_ptr, const int* __restrict__ seg_ptr, const int* __restrict__ other,
              const int* __restrict__ eid, const float* __restrict__ val, const float* __restrict__ dense, int n_nodes,
              int F, float* __restrict__ out) {
    const int lane = lane_id();
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_nodes; n += (gridDim.x * blockDim.x) >> 5) {
        const int beg = seg_ptr[node_ptr[n]], end = seg_ptr[node_ptr[n + 1]];
        for (int f0 = 0; f0 < F; f0 += 32) {
            const int f = f0 + lane;
            float a = 0.f;
            if (f < F)
                for (int p = beg; p < end; ++p) a = fmaf(val[eid[p]], dense[int64_t(other[p]) * F + f], a);
            if (f < F) out[int64_t(n) * F + f] = a;
        }
    }
}


// ------------------------------------------------------------------------------------------------ small glue
// out[c, r] = in[r, c]   (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
k_transpose(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int k = ty; k < 32; k += 8)
        if (r0 + k < rows && c0 + tx < cols) tile[k][tx] = in[size_t(r0 + k) * cols + c0 + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
        if (c0 + k < cols && r0 + tx < rows) out[size_t(c0 + k) * rows + r0 + tx] = tile[tx][k];
}

// FMEncoder glue (src/layers.py:541-547): x[i] = cat(e[i] / d_norm[i], h[i])  (mode 0)  or  e[i] / d_norm[i] + h[i]  (mode 1)
__global__ void k_drug_input_fwd(const float* __restrict__ e, const float* __restrict__ d_norm, const float* __restrict__ h,
                                 int n, int fe, int fh, int mode, float* __restrict__ out) {
    const int fo = mode == 0 ? fe + fh : fe;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(n) * fo) return;
    const int r = int(i / fo), c = int(i % fo);
    float v;
    if (mode == 0) v = c < fe ? e[size_t(r) * fe + c] / d_norm[r] : h[size_t(r) * fh + (c - fe)];
    else v = e[size_t(r) * fe + c] / d_norm[r] + h[size_t(r) * fh + c];
    out[i] = v;
}
__global__ void k_drug_input_bwd(const float* __restrict__ g, const float* __restrict__ d_norm, int n, int fe, int fh,
                                 int mode, float* __restrict__ d_e, float* __restrict__ d_h) {
    const int fo = mode == 0 ? fe + fh : fe;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(n) * fo) return;
    const int r = int(i / fo), c = int(i % fo);
    const float v = g[i];
    if (mode == 0) {
        if (c < fe) d_e[size_t(r) * fe + c] = v / d_norm[r];
        else d_h[size_t(r) * fh + (c - fe)] = v;
    } else {
        d_e[size_t(r) * fe + c] = v / d_norm[r];
        d_h[size_t(r) * fh + c] = v;
    }
}
__global__ void k_scale2(const float* __restrict__ a, int64_t na, const float* __restrict__ b, int64_t nb,
                         const float* __restrict__ scalar, float* __restrict__ oa, float* __restrict__ ob) {
    const float sc = *scalar;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < na) oa[i] = a[i] * sc;
    else if (i < na + nb) ob[i - na] = b[i - na] * sc;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_gemm(const float* a, const float* b, const float* bias, int64_t m, int64_t n, int64_t k, int trans_b, int relu,
              float* c, void* stream) {
    TIPB_CHECK_ARG(m >= 0 && n >= 0 && k >= 0 && ((a && b && c) || m * n == 0 || k == 0) && (c || m * n == 0), "gemm: bad argument");
    TIPB_CHECK_ARG(m < (int64_t(1) << 30) && n < (int64_t(1) << 24) && k < (int64_t(1) << 24), "gemm: shape too large for this kernel");
    return gemm_launch(a, b, bias, m, n, k, trans_b, relu, c, (cudaStream_t)stream);
}

constexpr int GEMM_TN_SLICES = 128;   // long-K products (K = 19,081 protein rows) need more CTAs than 32 slices give
size_t tipb_gemm_tn_workspace_bytes(int64_t m, int64_t n) { return size_t(GEMM_TN_SLICES) * size_t(m) * size_t(n) * 4 + 256; }

// out [m, n] = a^T b with a [k, m], b [k, n]: split over k into at most 32 slices that are summed in slice order
int tipb_gemm_tn(const float* a, const float* b, int64_t k, int64_t m, int64_t n, float* out, void* ws, size_t ws_bytes,
                 void* stream) {
    TIPB_CHECK_ARG(a && b && out && ws && k >= 0 && m > 0 && n > 0, "gemm_tn: bad argument");
    TIPB_CHECK_ARG(k < (int64_t(1) << 30) && m < (int64_t(1) << 24) && n < (int64_t(1) << 24), "gemm_tn: shape too large");
    TIPB_CHECK_ARG(ws_bytes >= tipb_gemm_tn_workspace_bytes(m, n), "gemm_tn: workspace too small");
    return atb_launch(a, b, (int)k, (int)m, (int)n, out, (float*)ws, (cudaStream_t)stream, GEMM_TN_SLICES);
}

size_t tipb_relu_grad_colsum_workspace_bytes(int64_t n) { return size_t(COLSUM_MAX_SLICES) * size_t(n) * 4 + 256; }

// g = gy where relu_out > 0 else 0 (relu_out NULL: no mask; g NULL: not written); d_bias[n] = column sums of g (NULL: none)
int tipb_relu_grad_colsum(const float* gy, const float* relu_out, int64_t m, int64_t n, float* g, float* d_bias, void* ws,
                          size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG((gy || m == 0) && m >= 0 && n > 0 && n < (int64_t(1) << 24) && (!relu_out || g), "relu_grad_colsum: bad argument");
    TIPB_CHECK_ARG(!d_bias || (ws && ws_bytes >= tipb_relu_grad_colsum_workspace_bytes(n)), "relu_grad_colsum: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (!g && !d_bias) return TIPB_OK;
    int n_slices = (int)ceil_div(m, 64);
    if (n_slices > COLSUM_MAX_SLICES) n_slices = COLSUM_MAX_SLICES;
    if (n_slices < 1) n_slices = 1;
    const int rows_per_slice = (int)ceil_div(m > 0 ? m : 1, n_slices);
    const dim3 grid((unsigned)ceil_div(n, 32), (unsigned)n_slices);
    k_mask_colsum<<<grid, 256, 0, s>>>(gy, relu_out, m, (int)n, rows_per_slice, g, d_bias ? (float*)ws : nullptr);
    if (d_bias) k_colsum_finish<<<(unsigned)ceil_div(n * 32, 256), 256, 0, s>>>((const float*)ws, n_slices, (int)n, d_bias);
    TIPB_CHECK_LAUNCH("relu_grad_colsum");
    return TIPB_OK;
}

int tipb_transpose(const float* in, int64_t rows, int64_t cols, float* out, void* stream) {
    TIPB_CHECK_ARG(in && out && rows >= 0 && cols >= 0 && rows < (int64_t(1) << 30) && cols < (int64_t(1) << 30), "transpose: bad argument");
    if (rows == 0 || cols == 0) return TIPB_OK;
    const dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    TIPB_CHECK_ARG(grid.y < 65536, "transpose: too many rows");
    k_transpose<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (int)rows, (int)cols, out);
    TIPB_CHECK_LAUNCH("transpose");
    return TIPB_OK;
}

int tipb_drug_input_fwd(const float* embed_out, const float* d_norm, const float* hier_out, int64_t n, int f_embed,
                        int f_hier, int mode, float* out, void* stream) {
    TIPB_CHECK_ARG(embed_out && d_norm && (hier_out || f_hier == 0) && out && (mode == 0 || (mode == 1 && f_embed == f_hier)), "drug_input_fwd: bad argument");
    const int64_t total = n * (mode == 0 ? f_embed + f_hier : f_embed);
    if (total > 0)
        k_drug_input_fwd<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(embed_out, d_norm, hier_out, (int)n, f_embed,
                                                                                          f_hier, mode, out);
    TIPB_CHECK_LAUNCH("drug_input_fwd");
    return TIPB_OK;
}

int tipb_drug_input_bwd(const float* grad, const float* d_norm, int64_t n, int f_embed, int f_hier, int mode,
                        float* d_embed_out, float* d_hier_out, void* stream) {
    TIPB_CHECK_ARG(grad && d_norm && d_embed_out && (d_hier_out || f_hier == 0) && (mode == 0 || (mode == 1 && f_embed == f_hier)), "drug_input_bwd: bad argument");
    const int64_t total = n * (mode == 0 ? f_embed + f_hier : f_embed);
    if (total > 0)
        k_drug_input_bwd<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(grad, d_norm, (int)n, f_embed, f_hier, mode,
                                                                                          d_embed_out, d_hier_out);
    TIPB_CHECK_LAUNCH("drug_input_bwd");
    return TIPB_OK;
}

int tipb_scale2(const float* a, int64_t n_a, const float* b, int64_t n_b, const float* scalar_dev, float* out_a,
                float* out_b, void* stream) {
    TIPB_CHECK_ARG(scalar_dev && (n_a == 0 || (a && out_a)) && (n_b == 0 || (b && out_b)), "scale2: bad argument");
    if (n_a + n_b > 0)
        k_scale2<<<(unsigned)ceil_div(n_a + n_b, 256), 256, 0, (cudaStream_t)stream>>>(a, n_a, b, n_b, scalar_dev, out_a, out_b);
    TIPB_CHECK_LAUNCH("scale2");
    return TIPB_OK;
}

int tipb_fill_zero(void* ptr, size_t bytes, void* stream) {
    if (bytes == 0) return TIPB_OK;
    TIPB_CHECK_ARG(ptr, "fill_zero: NULL argument");
    TIPB_CHECK_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
    return TIPB_OK;
}

int tipb_nn_decoder_fwd(const float* table_a, const float* table_b, const int64_t* edge_index, const int64_t* edge_type,
                        int64_t n_edges, int64_t n_nodes, int64_t n_rel, int apply_sigmoid, float* out, void* stream) {
    TIPB_CHECK_ARG(table_a && table_b && out && (n_edges == 0 || (edge_index && edge_type)), "nn_decoder_fwd: NULL argument");
    if (n_edges > 0)
        k_nn_decoder_fwd<<<(unsigned)ceil_div(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(
            table_a, table_b, edge_index, edge_type, n_edges, (int)n_nodes, (int)n_rel, apply_sigmoid, out);
    TIPB_CHECK_LAUNCH("nn_decoder_fwd");
    return TIPB_OK;
}

// d_table_a / d_table_b [n_nodes, n_rel] from the per-edge output gradient; plan_by_src / plan_by_dst: typed CSRs of
// the scored edges grouped by (row-0 endpoint, relation) and (row-1 endpoint, relation); g_ws: n_edges floats
int tipb_nn_decoder_bwd(const void* plan_by_src, const void* plan_by_dst, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                        const float* grad_out, const float* score, int apply_sigmoid, float* d_table_a, float* d_table_b,
                        float* g_ws, void* stream) {
    TIPB_CHECK_ARG(plan_by_src && plan_by_dst && grad_out && score && d_table_a && d_table_b && g_ws, "nn_decoder_bwd: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    TIPB_CHECK_CUDA(cudaMemsetAsync(d_table_a, 0, size_t(n_nodes) * n_rel * 4, s));
    TIPB_CHECK_CUDA(cudaMemsetAsync(d_table_b, 0, size_t(n_nodes) * n_rel * 4, s));
    if (n_edges == 0) return TIPB_OK;
    k_nn_decoder_gval<<<(unsigned)ceil_div(n_edges, 256), 256, 0, s>>>(grad_out, score, n_edges, apply_sigmoid, g_ws);
    const CsrView a = csr_view(plan_by_src, n_edges, n_nodes, n_rel), b = csr_view(plan_by_dst, n_edges, n_nodes, n_rel);
    const int grid = sm_count() * 8;
    k_seg_scalar_sum<<<grid, 256, 0, s>>>(a.seg_ptr, a.seg_node, a.seg_rel, a.eid, a.counts, g_ws, (int)n_rel, d_table_a);
    k_seg_scalar_sum<<<grid, 256, 0, s>>>(b.seg_ptr, b.seg_node, b.seg_rel, b.eid, b.counts, g_ws, (int)n_rel, d_table_b);
    TIPB_CHECK_LAUNCH("nn_decoder_bwd");
    return TIPB_OK;
}

// out [n_nodes, f] = S dense, S given as a typed CSR (n_rel = 1) over its COO entries: node = output row, other =
// column (row of `dense`), eid = position of the entry in `values`
int tipb_spmm_values(const void* plan, int64_t n_entries, int64_t n_nodes, const float* values, const float* dense, int f,
                     float* out, void* stream) {
    TIPB_CHECK_ARG(plan && values && dense && out && f > 0, "spmm_values: bad argument");
    const CsrView v = csr_view(plan, n_entries, n_nodes, 1);
    k_spmm_values<<<(unsigned)ceil_div(n_nodes * 32, 256), 256, 0, (cudaStream_t)stream>>>(v.node_ptr, v.seg_ptr, v.other, v.eid,
                                                                                          values, dense, (int)n_nodes, f, out);
    TIPB_CHECK_LAUNCH("spmm_values");
    return TIPB_OK;
}
}
