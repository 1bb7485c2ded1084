// The dense products around the R-GCN node contraction, batched over nodes.
//
// The tiled CUDA-core node kernels (rgcn_tiled.cuh) evaluate these per node inside the node's CTA, which makes every CTA
// re-read the whole basis tensor (256 KB for 32 x 64 x 32) from L2 with dependent loads: measured with the tensor-core
// node kernels, that part alone was 64 of 118 us (layer 1 forward) and 80 of 150 us (layer 1 backward).  Batched over
// nodes they are three small GEMMs in which a CTA shares one basis tile among 16-64 nodes:
//
//   forward    out[i,:]  = inv_deg_i * G[i,(b,f)] . basis[(b,f),:] + x_i root (+ bias) (+ ReLU)      K = B * F_in, split in slices
//   backward   Y[j,b,:]  = x_j . basis[b,:,:]                                                         (the d_att operand)
//              dX[j,:]   = sum_b Q[j,b,:] . basis[b,:,:]^T + geff_j root^T                            split over b in slices
//
// All three are one register-tiled kernel (4 nodes x 4 outputs per thread, operands staged per 32-wide K block); the
// root terms ride along as one extra slice.  Split partial sums are added in slice order by the finish kernel (deterministic).  Reference: the `torch.matmul(x_j,
// w)` with w = att @ basis of MyRGCNConv2.message and `x @ self.root` of .update (src/layers.py:166-188), reassociated.
#pragma once
#include "common.cuh"

namespace tipb {

constexpr int RD_THREADS = 256;
constexpr int RD_KB = 32;             // K block staged per trip (<= 32)
constexpr int RD_OUT_SLICES = 64;     // K = B * F_in split (forward): one 32-wide K block per CTA at B * F_in = 2048
constexpr int RD_DX_SLICES = 32;      // bases split (backward): one base per CTA

// One K-sliced product C_part[slice, i, :] = sum_{blocks of the slice} A[i, block] . B_block, register-tiled 4 nodes x 4
// outputs per thread.  A: row i at A + i * lda, block t covers columns [t * kb, (t + 1) * kb).  B_block t: element
// (k, o) at B + t * b_blk + k * sk + o * so  (sk / so select a plain or a transposed block).  `extra` is one more
// product of the same shape family written as slice n_slices (the root terms: x root, geff root^T).
struct RdProblem {
    const float* A;
    int lda;
    const float* B;
    int b_blk, sk, so;
    int kb;            // K per block (<= RD_KB)
    int n_blocks;      // blocks in total
};

template <int O>
__global__ void __launch_bounds__(RD_THREADS)
k_rd_partial(RdProblem main, RdProblem extra, int n_nodes, int blocks_per_slice, int n_slices, int c_col0_per_y, int ldc,
             float* __restrict__ C) {
    // blockIdx.y < n_slices: slice of `main`;  == n_slices: `extra`.  c_col0_per_y != 0: blockIdx.y is a batch index instead
    // (every y handles all blocks of its own B = main.B + y * main.b_blk * main.n_blocks ... see basis_y_launch)
    constexpr int O4 = O / 4, NQ = RD_THREADS / O4, NPC = NQ * 4;     // node quads / nodes per CTA
    __shared__ float4 As[RD_KB][NPC / 4 + 1];     // [k][node quad]: four consecutive nodes per float4
    __shared__ float4 Bs[RD_KB][O4];
    const int tid = threadIdx.x, o4 = tid % O4, nq = tid / O4;
    const int node0 = blockIdx.x * NPC, y = blockIdx.y;
    const bool batch = c_col0_per_y != 0;
    const bool is_extra = !batch && y == n_slices;
    const RdProblem P = is_extra ? extra : main;
    int t0 = 0, t1 = P.n_blocks;
    const float* Bbase = P.B;
    if (batch) Bbase += int64_t(y) * P.b_blk * P.n_blocks;
    else if (!is_extra) { t0 = y * blocks_per_slice; t1 = min(P.n_blocks, t0 + blocks_per_slice); }
    float4 acc[4] = {f4_zero(), f4_zero(), f4_zero(), f4_zero()};
    for (int t = t0; t < t1; ++t) {
        // A block -> As[k][node] (transposed while staged): 16-byte loads along k, all of a thread's loads issued first
        {
            constexpr int A_IT = NPC * (RD_KB / 4) / RD_THREADS;       // float4 loads per thread (NPC * 8 / 256)
            float4 va[A_IT];
#pragma unroll
            for (int u = 0; u < A_IT; ++u) {
                const int idx = tid + u * RD_THREADS, rr = idx / (RD_KB / 4), k4 = idx % (RD_KB / 4);
                va[u] = (k4 * 4 < P.kb && node0 + rr < n_nodes)
                            ? *reinterpret_cast<const float4*>(P.A + int64_t(node0 + rr) * P.lda + t * P.kb + k4 * 4) : f4_zero();
            }
#pragma unroll
            for (int u = 0; u < A_IT; ++u) {
                const int idx = tid + u * RD_THREADS, rr = idx / (RD_KB / 4), k4 = idx % (RD_KB / 4);
                float* dst = reinterpret_cast<float*>(&As[k4 * 4][rr >> 2]) + (rr & 3);
                constexpr int PITCH = (NPC / 4 + 1) * 4;
                dst[0] = va[u].x; dst[PITCH] = va[u].y; dst[2 * PITCH] = va[u].z; dst[3 * PITCH] = va[u].w;
            }
        }
        const float* Bt = Bbase + int64_t(t) * P.b_blk;
        if (P.so == 1) {      // rows of O contiguous floats: 16-byte loads
            constexpr int B_IT = (RD_KB * O4 + RD_THREADS - 1) / RD_THREADS;
#pragma unroll
            for (int u = 0; u < B_IT; ++u) {
                const int idx = tid + u * RD_THREADS;
                if (idx < RD_KB * O4) {
                    const int kk = idx / O4, q = idx % O4;
                    Bs[kk][q] = kk < P.kb ? *reinterpret_cast<const float4*>(Bt + kk * P.sk + q * 4) : f4_zero();
                }
            }
        } else {              // transposed block: k contiguous
            constexpr int B_IT = RD_KB * O / RD_THREADS;
            float vb[B_IT];
#pragma unroll
            for (int u = 0; u < B_IT; ++u) {
                const int idx = tid + u * RD_THREADS, o = idx / RD_KB, kk = idx % RD_KB;
                vb[u] = kk < P.kb ? Bt[kk * P.sk + o * P.so] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < B_IT; ++u) {
                const int idx = tid + u * RD_THREADS, o = idx / RD_KB, kk = idx % RD_KB;
                reinterpret_cast<float*>(&Bs[kk][0])[o] = vb[u];
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < RD_KB; ++kk) {
            const float4 a = As[kk][nq], b = Bs[kk][o4];
            acc[0] = f4_fma(a.x, b, acc[0]);
            acc[1] = f4_fma(a.y, b, acc[1]);
            acc[2] = f4_fma(a.z, b, acc[2]);
            acc[3] = f4_fma(a.w, b, acc[3]);
        }
        __syncthreads();
    }
    const int64_t plane = batch ? 0 : int64_t(y) * n_nodes * ldc;
    const int col0 = batch ? y * c_col0_per_y : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int node = node0 + nq * 4 + u;
        if (node < n_nodes) *reinterpret_cast<float4*>(C + plane + int64_t(node) * ldc + col0 + o4 * 4) = acc[u];
    }
}

// out[i, :] = act( scale_i * sum_{s < n_slices} part[s, i, :] + part[n_slices, i, :] + bias )      (float4 per thread)
__global__ void __launch_bounds__(RD_THREADS)
k_rd_finish(const float4* __restrict__ part, int n_slices, const float* __restrict__ scale, const float4* __restrict__ bias,
            int n_nodes, int w4, int relu, float4* __restrict__ out) {
    const int64_t gid = int64_t(blockIdx.x) * RD_THREADS + threadIdx.x;
    const int64_t plane = int64_t(n_nodes) * w4;
    if (gid >= plane) return;
    float4 a0 = f4_zero(), a1 = f4_zero();
    int k = 0;
    for (; k + 1 < n_slices; k += 2) {
        a0 = f4_add(a0, part[k * plane + gid]);
        a1 = f4_add(a1, part[(k + 1) * plane + gid]);
    }
    if (k < n_slices) a0 = f4_add(a0, part[k * plane + gid]);
    float4 s = f4_add(a0, a1);
    if (scale) {
        const float sc = scale[gid / w4];
        s.x *= sc; s.y *= sc; s.z *= sc; s.w *= sc;
    }
    s = f4_add(s, part[int64_t(n_slices) * plane + gid]);
    if (bias) s = f4_add(s, bias[gid % w4]);
    if (relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
    out[gid] = s;
}

static size_t rgcn_dense_fwd_ws_floats(int64_t n_nodes, int f_out) { return size_t(RD_OUT_SLICES + 1) * n_nodes * f_out; }
static size_t rgcn_dense_bwd_ws_floats(int64_t n_nodes, int f_in, int f_out, int n_bases) {
    return 2 * size_t(n_nodes) * n_bases * f_out + size_t(RD_DX_SLICES + 1) * n_nodes * f_in;     // Y, Q, dX partials
}

template <int O>
static void rd_launch(const RdProblem& m, const RdProblem& e, int n_nodes, int bps, int n_slices, int grid_y, int col0_per_y,
                      int ldc, float* C, cudaStream_t s) {
    constexpr int NPC = (RD_THREADS / (O / 4)) * 4;
    k_rd_partial<O><<<dim3((unsigned)ceil_div(n_nodes, NPC), grid_y), RD_THREADS, 0, s>>>(m, e, n_nodes, bps, n_slices, col0_per_y,
                                                                                         ldc, C);
}
#define RD_DISPATCH(W, ...)                                                                          \
    switch (W) {                                                                                      \
        case 128: rd_launch<128>(__VA_ARGS__); break;                                                 \
        case 64: rd_launch<64>(__VA_ARGS__); break;                                                   \
        case 32: rd_launch<32>(__VA_ARGS__); break;                                                   \
        case 16: rd_launch<16>(__VA_ARGS__); break;                                                   \
        default: set_last_error("rgcn_dense: width %d not a power of two in [16,128]", W); return TIPB_ERR_UNSUPPORTED; \
    }

// out = act( inv_deg * G basis + x root + bias )
static int basis_out_launch(const float* G, const float* basis, const float* inv_deg, const float* x, const float* root,
                            const float* bias, int n_nodes, int f_in, int f_out, int n_bases, int relu, float* partial,
                            float* out, cudaStream_t s) {
    const int K = n_bases * f_in;
    const int kb = K % RD_KB == 0 ? RD_KB : (K % 16 == 0 ? 16 : (K % 8 == 0 ? 8 : 4));
    const RdProblem m{G, K, basis, kb * f_out, f_out, 1, kb, K / kb};
    const int kbx = f_in % RD_KB == 0 ? RD_KB : (f_in % 16 == 0 ? 16 : (f_in % 8 == 0 ? 8 : 4));
    const RdProblem e{x, f_in, root, kbx * f_out, f_out, 1, kbx, f_in / kbx};
    const int bps = (int)ceil_div(m.n_blocks, RD_OUT_SLICES), n_slices = (int)ceil_div(m.n_blocks, bps);
    RD_DISPATCH(f_out, m, e, n_nodes, bps, n_slices, n_slices + 1, 0, f_out, partial, s)
    k_rd_finish<<<(unsigned)ceil_div(int64_t(n_nodes) * f_out / 4, RD_THREADS), RD_THREADS, 0, s>>>(
        (const float4*)partial, n_slices, inv_deg, (const float4*)bias, n_nodes, f_out / 4, relu, (float4*)out);
    TIPB_CHECK_LAUNCH("basis_out");
    return TIPB_OK;
}

// Y[j, b, :] = x_j basis[b]          (batch over b: blockIdx.y)
static int basis_y_launch(const float* x, const float* basis, int n_nodes, int f_in, int f_out, int n_bases, float* Y,
                          cudaStream_t s) {
    const int kb = f_in % RD_KB == 0 ? RD_KB : (f_in % 16 == 0 ? 16 : (f_in % 8 == 0 ? 8 : 4));
    const RdProblem m{x, f_in, basis, kb * f_out, f_out, 1, kb, f_in / kb};
    RD_DISPATCH(f_out, m, m, n_nodes, 0, 0, n_bases, f_out, n_bases * f_out, Y, s)
    TIPB_CHECK_LAUNCH("basis_y");
    return TIPB_OK;
}

// dX = sum_b Q[:, b, :] basis[b]^T + geff root^T
static int basis_dx_launch(const float* Q, const float* basis, const float* geff, const float* root, int n_nodes, int f_in,
                           int f_out, int n_bases, float* partial, float* d_x, cudaStream_t s) {
    if (f_out > RD_KB) { set_last_error("basis_dx: f_out=%d > %d", f_out, RD_KB); return TIPB_ERR_UNSUPPORTED; }
    // block t = base b: A columns [b * f_out, (b + 1) * f_out), B_block(k = o, f) = basis[b, f, o]: transposed block
    const RdProblem m{Q, n_bases * f_out, basis, f_in * f_out, 1, f_out, f_out, n_bases};
    const RdProblem e{geff, f_out, root, 0, 1, f_out, f_out, 1};
    const int bps = (int)ceil_div(n_bases, RD_DX_SLICES), n_slices = (int)ceil_div(n_bases, bps);
    RD_DISPATCH(f_in, m, e, n_nodes, bps, n_slices, n_slices + 1, 0, f_in, partial, s)
    k_rd_finish<<<(unsigned)ceil_div(int64_t(n_nodes) * f_in / 4, RD_THREADS), RD_THREADS, 0, s>>>(
        (const float4*)partial, n_slices, nullptr, nullptr, n_nodes, f_in / 4, 0, (float4*)d_x);
    TIPB_CHECK_LAUNCH("basis_dx");
    return TIPB_OK;
}
#undef RD_DISPATCH

}  // namespace tipb
