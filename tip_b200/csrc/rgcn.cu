// Basis-decomposed R-GCN layer, forward and backward (replaces MyRGCNConv / MyRGCNConv2,
// reference src/layers.py:21-99, 102-193, and what torch autograd derives from them).
//
//   out_i = 1/max(deg_i,1) * sum_{e=(j->i), r=type(e)} x_j W_r + x_i root,   W_r = sum_b att[r,b] basis[b]
//
// Reassociated so that no E x F tensor is ever materialised (the reference builds x_j [E,F_in],
// the concatenated messages [E,F_out] and, in backward, one zero-filled E x F_in buffer per relation):
//
//   forward   H[s,:]   = sum_{e in segment s=(i,r)} x_j                     seg_aggregate  (edge pass, payload F_in)
//             G[i,b,:] = sum_{s in node i} att[r_s,b] H[s,:]               k_rgcn_node_fwd (one CTA per node,
//             out_i    = inv_deg_i * <G[i], basis> + x_i root                 a [B x S_i] x [S_i x F_in] product)
//   backward  ghat_i   = inv_deg_i * gout_i  (ReLU mask folded in)
//             T[s,:]   = sum_{e in segment s=(j,r) of the by-source CSR} ghat_i   seg_aggregate (payload F_out)
//             Q[j,b,:] = sum_s att[r_s,b] T[s,:] ;  dX_j = <Q[j], basis^T> + gout_j root^T
//             d_att[r,b] = sum_{s: r_s=r} <T[s], Y[j_s,b,:]>,  Y[j,b,:] = x_j basis[b]   (per-segment dots, then a
//                                                                                   relation-major reduction)
//             d_basis[b,f,o] = sum_i G[i,b,f] ghat_i[o] ;  d_root = X^T gout          (k_atb, split-K, fixed order)
// Every sum has a fixed order: no floating-point atomics anywhere.
#include "common.cuh"
#include "seg_aggregate.cuh"
#include "reduce.cuh"
#include "rgcn_tiled.cuh"
#include "rgcn_tc.cuh"
#include "rgcn_dense.cuh"

namespace tipb {

constexpr int NODE_THREADS = 256;
constexpr int CH = 16;  // segments staged per pipeline step

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float dot4(float4 a, float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// stage `cn` segment rows (payload [*, tf float4]) and their relation's att rows into shared memory
__device__ __forceinline__ void stage_chunk(float4* sP, float* sA, const float4* __restrict__ payload,
                                            const float* __restrict__ att, const int* __restrict__ seg_rel, int s0,
                                            int cn, int tf, int n_bases) {
    for (int i = threadIdx.x; i < cn * tf; i += NODE_THREADS) {
        int c = i / tf, q = i - c * tf;
        cp_async16(&sP[c * tf + q], &payload[int64_t(s0 + c) * tf + q]);
    }
    for (int i = threadIdx.x; i < cn * n_bases; i += NODE_THREADS) {
        int c = i / n_bases, b = i - c * n_bases;
        cp_async4(&sA[c * n_bases + b], &att[int64_t(seg_rel[s0 + c]) * n_bases + b]);
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------------
// forward node kernel: one CTA per target node i
// smem: [2][CH*tf] float4 | [2][CH*B] float | Gs [B*f_in] | xs [f_in] | red [256]
template <int NB>
__global__ void __launch_bounds__(NODE_THREADS)
k_rgcn_node_fwd(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float* __restrict__ inv_deg,
                const float4* __restrict__ H, const float* __restrict__ att, const float* __restrict__ basis,
                const float* __restrict__ root, const float* __restrict__ bias, const float* __restrict__ x,
                int f_in, int f_out, int n_bases, int relu, float* __restrict__ out, float4* __restrict__ g_saved) {
    extern __shared__ float4 smem4[];
    const int tf = f_in >> 2;
    const int tb = NODE_THREADS / tf;
    const int tx = threadIdx.x % tf, ty = threadIdx.x / tf;
    float4* sH = smem4;                                   // 2 * CH * tf
    float* sA = reinterpret_cast<float*>(sH + 2 * CH * tf);  // 2 * CH * B
    float* Gs = sA + 2 * CH * n_bases;                    // B * f_in   (16-byte aligned: CH*B*2 floats, CH=16)
    float* xs = Gs + n_bases * f_in;                      // f_in
    float* red = xs + f_in;                               // NODE_THREADS

    const int i = blockIdx.x;
    const int sb = node_ptr[i], se = node_ptr[i + 1];
    const int n_chunks = (se - sb + CH - 1) / CH;

    for (int f = threadIdx.x; f < f_in; f += NODE_THREADS) xs[f] = x[int64_t(i) * f_in + f];

    float4 acc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[k] = f4_zero();

    if (n_chunks > 0) stage_chunk(sH, sA, H, att, seg_rel, sb, min(CH, se - sb), tf, n_bases);
    for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        if (c + 1 < n_chunks) {
            const int s0 = sb + (c + 1) * CH;
            stage_chunk(sH + (st ^ 1) * CH * tf, sA + (st ^ 1) * CH * n_bases, H, att, seg_rel, s0, min(CH, se - s0), tf,
                        n_bases);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int cn = min(CH, se - (sb + c * CH));
        const float4* h4 = sH + st * CH * tf;
        const float* a = sA + st * CH * n_bases;
        for (int q = 0; q < cn; ++q) {
            const float4 h = h4[q * tf + tx];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int b = ty + k * tb;
                if (b < n_bases) acc[k] = f4_fma(a[q * n_bases + b], h, acc[k]);
            }
        }
        __syncthreads();  // stage st is refilled two iterations from now; everyone must be done reading it
    }

    // G tile -> global (saved for d_basis) and shared (for the basis contraction)
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const int b = ty + k * tb;
        if (b < n_bases) {
            g_saved[(int64_t(i) * n_bases + b) * tf + tx] = acc[k];
            reinterpret_cast<float4*>(Gs)[b * tf + tx] = acc[k];
        }
    }
    __syncthreads();

    // out_i[o] = inv_deg_i * sum_{b,f} G[b,f] basis[b,f,o] + sum_f x_i[f] root[f,o]
    const int o = threadIdx.x % f_out, q = threadIdx.x / f_out, nq = NODE_THREADS / f_out;
    float sum = 0.f;
    const int bf_total = n_bases * f_in;
    for (int bf = q; bf < bf_total; bf += nq) sum = fmaf(Gs[bf], basis[int64_t(bf) * f_out + o], sum);
    sum *= inv_deg[i];
    for (int f = q; f < f_in; f += nq) sum = fmaf(xs[f], root[f * f_out + o], sum);
    red[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x < f_out) {
        float tot = 0.f;
        for (int qq = 0; qq < nq; ++qq) tot += red[qq * f_out + threadIdx.x];
        if (bias) tot += bias[threadIdx.x];
        if (relu) tot = fmaxf(tot, 0.f);
        out[int64_t(i) * f_out + threadIdx.x] = tot;
    }
}

// ------------------------------------------------------------------------------------------------
// backward node kernel: one CTA per source node j (segments of the by-source CSR)
// smem: [2][CH*tfo] float4 | [2][CH*B] float | Ds [CH*B] | Qs [B*f_out] | xs [f_in] | gs [f_out]
template <int NB>
__global__ void __launch_bounds__(NODE_THREADS)
k_rgcn_node_bwd(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float4* __restrict__ T,
                const float* __restrict__ att, const float4* __restrict__ basis4, const float4* __restrict__ root4,
                const float* __restrict__ x, const float* __restrict__ geff, int f_in, int f_out, int n_bases,
                float* __restrict__ datt_seg, float* __restrict__ d_x) {
    extern __shared__ float4 smem4[];
    const int tfo = f_out >> 2;
    const int tb = NODE_THREADS / tfo;
    const int tx = threadIdx.x % tfo, ty = threadIdx.x / tfo;
    float4* sT = smem4;
    float* sA = reinterpret_cast<float*>(sT + 2 * CH * tfo);
    float* Ds = sA + 2 * CH * n_bases;
    float* Qs = Ds + CH * n_bases;       // offset (3*CH*B) floats: multiple of 4 since CH = 16
    float* xs = Qs + n_bases * f_out;
    float* gs = xs + f_in;               // f_in % 4 == 0 keeps 16-byte alignment

    const int j = blockIdx.x;
    const int sb = node_ptr[j], se = node_ptr[j + 1];
    const int n_chunks = (se - sb + CH - 1) / CH;

    for (int f = threadIdx.x; f < f_in; f += NODE_THREADS) xs[f] = x[int64_t(j) * f_in + f];
    for (int o = threadIdx.x; o < f_out; o += NODE_THREADS) gs[o] = geff[int64_t(j) * f_out + o];
    if (n_chunks > 0) stage_chunk(sT, sA, T, att, seg_rel, sb, min(CH, se - sb), tfo, n_bases);
    __syncthreads();

    // Y[j,b,o4] = sum_f x_j[f] basis[b,f,o4]
    float4 y[NB], qacc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        y[k] = f4_zero();
        qacc[k] = f4_zero();
        const int b = ty + k * tb;
        if (b < n_bases && n_chunks > 0) {
            const float4* bp = basis4 + (int64_t(b) * f_in) * tfo + tx;
            for (int f = 0; f < f_in; ++f) y[k] = f4_fma(xs[f], bp[int64_t(f) * tfo], y[k]);
        }
    }

    for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        const int s0 = sb + c * CH;
        if (c + 1 < n_chunks) {
            const int s1 = s0 + CH;
            stage_chunk(sT + (st ^ 1) * CH * tfo, sA + (st ^ 1) * CH * n_bases, T, att, seg_rel, s1, min(CH, se - s1),
                        tfo, n_bases);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int cn = min(CH, se - s0);
        const float4* t4 = sT + st * CH * tfo;
        const float* a = sA + st * CH * n_bases;
        for (int q = 0; q < cn; ++q) {
            const float4 t = t4[q * tfo + tx];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int b = ty + k * tb;
                const bool ok = b < n_bases;
                if (ok) qacc[k] = f4_fma(a[q * n_bases + b], t, qacc[k]);
                float d = ok ? dot4(t, y[k]) : 0.f;
                for (int off = tfo >> 1; off > 0; off >>= 1) d += __shfl_xor_sync(FULL, d, off);
                if (ok && tx == 0) Ds[q * n_bases + b] = d;
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < cn * n_bases; idx += NODE_THREADS)
            datt_seg[int64_t(s0) * n_bases + idx] = Ds[idx];
        // the top-of-loop barrier of the next iteration orders these reads before Ds is rewritten
    }

#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const int b = ty + k * tb;
        if (b < n_bases) reinterpret_cast<float4*>(Qs)[b * tfo + tx] = qacc[k];
    }
    __syncthreads();

    // dX_j[f] = sum_{b,o} Q[b,o] basis[b,f,o] + sum_o geff_j[o] root[f,o]
    const float4* Q4 = reinterpret_cast<const float4*>(Qs);
    const float4 g4 = reinterpret_cast<const float4*>(gs)[tx];
    for (int f = ty; f < f_in; f += tb) {
        float sum = dot4(g4, root4[f * tfo + tx]);
        for (int b = 0; b < n_bases; ++b) sum += dot4(Q4[b * tfo + tx], basis4[(int64_t(b) * f_in + f) * tfo + tx]);
        for (int off = tfo >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);
        if (tx == 0) d_x[int64_t(j) * f_in + f] = sum;
    }
}

// gout -> (geff = gout * relu'(out), ghat = inv_deg * geff)
__global__ void k_rgcn_grad_prep(const float* __restrict__ gout, const float* __restrict__ out_ref,
                                 const float* __restrict__ inv_deg, int64_t n, int f_out, float* __restrict__ geff,
                                 float* __restrict__ ghat) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * f_out) return;
    float g = gout[i];
    if (out_ref && !(out_ref[i] > 0.f)) g = 0.f;
    geff[i] = g;
    ghat[i] = g * inv_deg[i / f_out];
}

// out_partial[ks][m][n] = sum_{i in K-slice ks} A[i][m] * scale[i] * Bm[i][n]      (A^T B, split over K)
constexpr int ATB_ROWS = 32, ATB_MAXQ = 16;
__global__ void __launch_bounds__(256)
k_atb(const float* __restrict__ A, const float* __restrict__ Bm, int K, int M, int N, int k_slices,
      float* __restrict__ partial) {
    const int o = threadIdx.x % N, rg = threadIdx.x / N, n_rg = 256 / N;
    const int row0 = blockIdx.x * ATB_ROWS;
    const int ks = blockIdx.y;
    const int kc = (K + k_slices - 1) / k_slices;
    const int kb = ks * kc, ke = min(K, kb + kc);
    float acc[ATB_MAXQ];
#pragma unroll
    for (int q = 0; q < ATB_MAXQ; ++q) acc[q] = 0.f;
    for (int i = kb; i < ke; ++i) {
        const float bv = Bm[int64_t(i) * N + o];
        const float* arow = A + int64_t(i) * M + row0;
#pragma unroll
        for (int q = 0; q < ATB_MAXQ; ++q) {
            const int rr = rg + q * n_rg;
            if (rr < ATB_ROWS && row0 + rr < M) acc[q] = fmaf(arow[rr], bv, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < ATB_MAXQ; ++q) {
        const int rr = rg + q * n_rg;
        if (rr < ATB_ROWS && row0 + rr < M) partial[(int64_t(ks) * M + row0 + rr) * N + o] = acc[q];
    }
}

__global__ void k_sum_slices(const float* __restrict__ partial, int64_t n, int k_slices, float* __restrict__ out) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < k_slices; ++k) s += partial[int64_t(k) * n + i];
    out[i] = s;
}

// shared-memory tiled variant: CTA tile 64 (M) x N, K chunk 32; each thread owns RPT rows x 4 columns
// (RPT = N / 16), reads one float4 of B and RPT scalars of A per k and issues RPT FFMA2 pairs.
constexpr int ATBT_ROWS = 64, ATBT_KC = 32;
template <int N>
__global__ void __launch_bounds__(256)
k_atb_tiled(const float* __restrict__ A, const float* __restrict__ Bm, int K, int M, int k_slices,
            float* __restrict__ partial) {
    constexpr int RPT = N / 16, CG = N / 4;
    __shared__ float As[ATBT_KC][ATBT_ROWS];
    __shared__ float4 Bs[ATBT_KC][CG];
    const int tid = threadIdx.x;
    const int c4 = tid % CG, rg = tid / CG;
    const int row0 = blockIdx.x * ATBT_ROWS;
    const int ks = blockIdx.y;
    const int kc = (K + k_slices - 1) / k_slices;
    const int kb = ks * kc, ke = min(K, kb + kc);
    float4 acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i] = f4_zero();
    for (int k0 = kb; k0 < ke; k0 += ATBT_KC) {
        const int kn = min(ATBT_KC, ke - k0);
        for (int i = tid; i < ATBT_KC * ATBT_ROWS; i += 256) {
            const int kk = i / ATBT_ROWS, r = i % ATBT_ROWS;
            As[kk][r] = (kk < kn && row0 + r < M) ? A[int64_t(k0 + kk) * M + row0 + r] : 0.f;
        }
        for (int i = tid; i < ATBT_KC * CG; i += 256) {
            const int kk = i / CG, c = i % CG;
            Bs[kk][c] = kk < kn ? reinterpret_cast<const float4*>(Bm)[int64_t(k0 + kk) * CG + c] : f4_zero();
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < ATBT_KC; ++kk) {
            const float4 b = Bs[kk][c4];
#pragma unroll
            for (int i = 0; i < RPT; ++i) acc[i] = f4_fma(As[kk][rg * RPT + i], b, acc[i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int row = row0 + rg * RPT + i;
        if (row < M) reinterpret_cast<float4*>(partial)[(int64_t(ks) * M + row) * CG + c4] = acc[i];
    }
}

int atb_launch(const float* A, const float* Bm, int K, int M, int N, float* out, float* partial_ws, cudaStream_t s,
               int slice_cap) {
    if (N == 16 || N == 32 || N == 64 || N == 128) {
        const int tiles = (int)ceil_div(M, ATBT_ROWS);
        int k_slices = (int)ceil_div(2 * sm_count(), tiles);
        if (k_slices > slice_cap) k_slices = slice_cap;
        const int max_slices = (int)ceil_div(K > 0 ? K : 1, ATBT_KC);
        if (k_slices > max_slices) k_slices = max_slices;
        const dim3 grid(tiles, k_slices);
        if (N == 16) k_atb_tiled<16><<<grid, 256, 0, s>>>(A, Bm, K, M, k_slices, partial_ws);
        else if (N == 32) k_atb_tiled<32><<<grid, 256, 0, s>>>(A, Bm, K, M, k_slices, partial_ws);
        else if (N == 64) k_atb_tiled<64><<<grid, 256, 0, s>>>(A, Bm, K, M, k_slices, partial_ws);
        else k_atb_tiled<128><<<grid, 256, 0, s>>>(A, Bm, K, M, k_slices, partial_ws);
        k_sum_slices<<<(unsigned)ceil_div(int64_t(M) * N, 256), 256, 0, s>>>(partial_ws, int64_t(M) * N, k_slices, out);
        TIPB_CHECK_LAUNCH("atb_tiled");
        return TIPB_OK;
    }
    // enough CTAs to cover the machine, but never more slices than rows of K
    int tiles = (int)ceil_div(M, ATB_ROWS);
    int k_slices = (int)ceil_div(2 * sm_count(), tiles);
    if (k_slices > slice_cap) k_slices = slice_cap;
    if (k_slices > K) k_slices = K > 0 ? K : 1;
    k_atb<<<dim3(tiles, k_slices), 256, 0, s>>>(A, Bm, K, M, N, k_slices, partial_ws);
    k_sum_slices<<<(unsigned)ceil_div(int64_t(M) * N, 256), 256, 0, s>>>(partial_ws, int64_t(M) * N, k_slices, out);
    TIPB_CHECK_LAUNCH("atb");
    return TIPB_OK;
}
int atb_launch(const float* A, const float* Bm, int K, int M, int N, float* out, float* partial_ws, cudaStream_t s) {
    return atb_launch(A, Bm, K, M, N, out, partial_ws, s, 32);
}
size_t atb_ws_floats(int M, int N) { return size_t(32) * M * N; }

static bool pow2_in(int v, int lo, int hi) { return v >= lo && v <= hi && (v & (v - 1)) == 0; }

static int pick_nb(int n_bases, int tb) {
    int nb = (n_bases + tb - 1) / tb;
    if (nb <= 1) return 1;
    if (nb <= 2) return 2;
    if (nb <= 4) return 4;
    if (nb <= 8) return 8;
    return -1;
}

static size_t node_fwd_smem(int f_in, int n_bases) {
    return size_t(2 * CH * f_in + 2 * CH * n_bases + n_bases * f_in + f_in + NODE_THREADS) * sizeof(float);
}
static size_t node_bwd_smem(int f_in, int f_out, int n_bases) {
    return size_t(2 * CH * f_out + 3 * CH * n_bases + n_bases * f_out + f_in + f_out) * sizeof(float);
}

static int check_dims(const char* who, int f_in, int f_out, int n_bases) {
    if (!pow2_in(f_in, 4, 128) || !pow2_in(f_out, 4, 128)) {
        set_last_error("%s: f_in=%d / f_out=%d must be powers of two in [4,128] (the Python layer pads)", who, f_in, f_out);
        return TIPB_ERR_UNSUPPORTED;
    }
    if (n_bases < 1 || pick_nb(n_bases, NODE_THREADS / (f_in / 4)) < 0 || pick_nb(n_bases, NODE_THREADS / (f_out / 4)) < 0 ||
        node_fwd_smem(f_in, n_bases) > size_t(max_smem_optin()) || node_bwd_smem(f_in, f_out, n_bases) > size_t(max_smem_optin())) {
        set_last_error("%s: n_bases=%d too large for f_in=%d f_out=%d", who, n_bases, f_in, f_out);
        return TIPB_ERR_UNSUPPORTED;
    }
    return TIPB_OK;
}

struct FwdWs { float* H; };
struct BwdWs { float *T, *datt_seg, *geff, *ghat, *partial; };

static size_t fwd_ws_bytes(int64_t seg_cap, int64_t n_nodes, int f_in, int f_out) {
    return (size_t(seg_cap) * f_in + rgcn_dense_fwd_ws_floats(n_nodes, f_out)) * 4 + 1024;
}
static size_t bwd_ws_bytes(int64_t seg_cap, int64_t n_nodes, int f_in, int f_out, int n_bases) {
    size_t m = size_t(n_bases) * f_in > size_t(f_in) ? size_t(n_bases) * f_in : size_t(f_in);
    return (size_t(seg_cap) * (f_out + n_bases) + 2 * size_t(n_nodes) * f_out + atb_ws_floats((int)m, f_out) +
            rgcn_dense_bwd_ws_floats(n_nodes, f_in, f_out, n_bases)) * 4 + 4096;
}

constexpr int64_t RGCN_TC_MIN_REL = 256;     // relations (= upper bound of a node's segments) from which tcgen05 pays off
static int rgcn_tc_dbg() {
    static const int v = [] { const char* e = getenv("TIPB_RGCN_TC_DBG"); return e ? atoi(e) : 0; }();
    return v;
}
int* rgcn_tc_error_flag() {
    static int* ptr = nullptr;
    if (!ptr) cudaGetSymbolAddress(reinterpret_cast<void**>(&ptr), g_rgcn_tc_error);
    return ptr;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_seg_aggregate(const void* plan, int64_t n_entries, int64_t n_nodes, int64_t n_rel, const float* feat,
                       int64_t n_feat_rows, int f, float* out, void* stream) {
    TIPB_CHECK_ARG(plan && feat && out, "seg_aggregate: NULL argument");
    CsrView v = csr_view(plan, n_entries, n_nodes, n_rel);
    return seg_aggregate_launch(v, feat, nullptr, nullptr, (int)n_feat_rows, f, out, (cudaStream_t)stream);
}

size_t tipb_rgcn_workspace_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel, int f_in, int f_out, int n_bases) {
    int64_t cap = n_nodes * n_rel;
    int64_t seg_cap = n_entries < cap ? n_entries : cap;
    if (seg_cap < 1) seg_cap = 1;
    size_t a = fwd_ws_bytes(seg_cap, n_nodes, f_in, f_out), b = bwd_ws_bytes(seg_cap, n_nodes, f_in, f_out, n_bases);
    return a > b ? a : b;
}

// 0 = every tensor-core node kernel so far completed its barrier protocol (one blocking device read)
int tipb_rgcn_tc_status(void) {
    int v = 0;
    if (cudaMemcpy(&v, rgcn_tc_error_flag(), sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

int tipb_rgcn_fwd(const void* plan_by_dst, int64_t n_entries, int64_t n_nodes, int64_t n_rel, const float* x,
                  const float* basis, const float* att, const float* root, const float* bias, int f_in, int f_out,
                  int n_bases, int relu_out, float* out, float* g_saved, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_dims("rgcn_fwd", f_in, f_out, n_bases);
    if (rc) return rc;
    TIPB_CHECK_ARG(plan_by_dst && x && basis && att && root && out && g_saved && ws, "rgcn_fwd: NULL argument");
    CsrView v = csr_view(plan_by_dst, n_entries, n_nodes, n_rel);
    TIPB_CHECK_ARG(ws_bytes >= fwd_ws_bytes(v.seg_cap, n_nodes, f_in, f_out), "rgcn_fwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    Carver c(ws);
    float* H = c.take<float>(size_t(v.seg_cap) * f_in);
    float* out_partial = c.take<float>(rgcn_dense_fwd_ws_floats(n_nodes, f_out));
    if ((rc = seg_aggregate_launch(v, x, nullptr, nullptr, (int)n_nodes, f_in, H, s))) return rc;

    // tensor-core node contraction (rgcn_tc.cuh) for the shapes of the TIP / D-D nets; TIPB_RGCN_TC=0 selects the
    // CUDA-core kernels below (measurements, and the shapes the tcgen05 form does not cover)
    // (a node of a relation shard has few segments -- 861 / 8 relations on 8 GPUs: the per-node set-up of the tcgen05
    // kernel then outweighs its main loop, and the tiled kernel with its five CTAs per SM is the faster one)
    static const bool tc_env = [] { const char* e = getenv("TIPB_RGCN_TC"); return !(e && e[0] == '0'); }();
    const bool use_tc = tc_env && n_rel >= RGCN_TC_MIN_REL;
#define TC_FWD(FV, NBV)                                                                                            \
    {                                                                                                              \
        auto kern = k_rgcn_node_fwd_tc<FV, NBV>;                                                                   \
        const size_t sm = RtLayout<FV, NBV>::BYTES;                                                                \
        if ((rc = ensure_dyn_smem((const void*)kern, sm))) return rc;                                              \
        kern<<<(unsigned)n_nodes, RT_THREADS, sm, s>>>(v.node_ptr, v.seg_rel, H, att, g_saved, rgcn_tc_error_flag(), \
                                                       rgcn_tc_dbg());                                             \
        TIPB_CHECK_LAUNCH("rgcn_node_fwd_tc");                                                                     \
        return basis_out_launch(g_saved, basis, v.inv_deg, x, root, bias, (int)n_nodes, f_in, f_out, n_bases,      \
                                relu_out, out_partial, out, s);                                                    \
    }
    if (use_tc && (f_out == 16 || f_out == 32 || f_out == 64 || f_out == 128)) {
        if (n_bases == 32) {
            if (f_in == 64) TC_FWD(64, 32)
            if (f_in == 32) TC_FWD(32, 32)
        } else if (n_bases == 16) {
            if (f_in == 64) TC_FWD(64, 16)
            if (f_in == 32) TC_FWD(32, 16)
        }
    }
#undef TC_FWD

    // register-tiled kernel for the common shapes
#define TILED_FWD(TFV, NBQV)                                                                                       \
    {                                                                                                              \
        auto kern = k_rgcn_node_fwd_tiled<TFV, NBQV>;                                                              \
        const size_t sm = node_fwd_tiled_smem<TFV, NBQV>();                                                        \
        if ((rc = ensure_dyn_smem((const void*)kern, sm))) return rc;                                              \
        kern<<<(unsigned)n_nodes, TILED_THREADS, sm, s>>>(v.node_ptr, v.seg_rel, v.inv_deg, (const float4*)H,      \
                                                          (const float4*)att, basis, root, bias, x, f_out, relu_out, \
                                                          out, (float4*)g_saved);                                  \
        TIPB_CHECK_LAUNCH("rgcn_node_fwd_tiled");                                                                  \
        return TIPB_OK;                                                                                            \
    }
    if (f_out <= TILED_THREADS) {
        if (n_bases == 32) {
            if (f_in == 64) TILED_FWD(16, 8)
            if (f_in == 32) TILED_FWD(8, 8)
            if (f_in == 16) TILED_FWD(4, 8)
        } else if (n_bases == 16) {
            if (f_in == 64) TILED_FWD(16, 4)
            if (f_in == 32) TILED_FWD(8, 4)
            if (f_in == 16) TILED_FWD(4, 4)
        }
    }
#undef TILED_FWD

    const int nb = pick_nb(n_bases, NODE_THREADS / (f_in / 4));
    const size_t smem = node_fwd_smem(f_in, n_bases);
#define LAUNCH_FWD(NBV)                                                                                           \
    {                                                                                                              \
        auto kern = k_rgcn_node_fwd<NBV>;                                                                          \
        if ((rc = ensure_dyn_smem((const void*)kern, smem))) return rc; \
        kern<<<(unsigned)n_nodes, NODE_THREADS, smem, s>>>(v.node_ptr, v.seg_rel, v.inv_deg, (const float4*)H, att, basis, \
                                                           root, bias, x, f_in, f_out, n_bases, relu_out, out,    \
                                                           (float4*)g_saved);                                      \
    }
    switch (nb) {
        case 1: LAUNCH_FWD(1) break;
        case 2: LAUNCH_FWD(2) break;
        case 4: LAUNCH_FWD(4) break;
        default: LAUNCH_FWD(8) break;
    }
#undef LAUNCH_FWD
    TIPB_CHECK_LAUNCH("rgcn_node_fwd");
    return TIPB_OK;
}

int tipb_rgcn_bwd(const void* plan_by_src, int64_t n_entries, int64_t n_nodes, int64_t n_rel, const float* inv_deg_dst,
                  const float* x, const float* basis, const float* att, const float* root, const float* g_saved,
                  const float* grad_out, const float* out_for_relu, int f_in, int f_out, int n_bases, float* d_x,
                  float* d_basis, float* d_att, float* d_root, float* d_bias, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_dims("rgcn_bwd", f_in, f_out, n_bases);
    if (rc) return rc;
    TIPB_CHECK_ARG(plan_by_src && inv_deg_dst && x && basis && att && root && g_saved && grad_out && d_x && d_basis &&
                       d_att && d_root && ws, "rgcn_bwd: NULL argument");
    TIPB_CHECK_ARG(d_bias == nullptr, "rgcn_bwd: d_bias is reduced by the caller (bias is unused on the TIP path)");
    CsrView v = csr_view(plan_by_src, n_entries, n_nodes, n_rel);
    TIPB_CHECK_ARG(ws_bytes >= bwd_ws_bytes(v.seg_cap, n_nodes, f_in, f_out, n_bases), "rgcn_bwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    Carver c(ws);
    float* T = c.take<float>(size_t(v.seg_cap) * f_out);
    float* datt_seg = c.take<float>(size_t(v.seg_cap) * n_bases);
    float* geff = c.take<float>(size_t(n_nodes) * f_out);
    float* ghat = c.take<float>(size_t(n_nodes) * f_out);
    const int m_basis = n_bases * f_in;
    float* partial = c.take<float>(atb_ws_floats(m_basis > f_in ? m_basis : f_in, f_out));
    float* Ybuf = c.take<float>(size_t(n_nodes) * n_bases * f_out);
    float* Qbuf = c.take<float>(size_t(n_nodes) * n_bases * f_out);
    float* dx_partial = c.take<float>(size_t(RD_DX_SLICES + 1) * n_nodes * f_in);

    k_rgcn_grad_prep<<<(unsigned)ceil_div(n_nodes * f_out, 256), 256, 0, s>>>(grad_out, out_for_relu, inv_deg_dst, n_nodes,
                                                                              f_out, geff, ghat);
    if ((rc = seg_aggregate_launch(v, ghat, nullptr, nullptr, (int)n_nodes, f_out, T, s))) return rc;

    const int nb = pick_nb(n_bases, NODE_THREADS / (f_out / 4));
    const size_t smem = node_bwd_smem(f_in, f_out, n_bases);
#define LAUNCH_BWD(NBV)                                                                                            \
    {                                                                                                              \
        auto kern = k_rgcn_node_bwd<NBV>;                                                                          \
        if ((rc = ensure_dyn_smem((const void*)kern, smem))) return rc; \
        kern<<<(unsigned)n_nodes, NODE_THREADS, smem, s>>>(v.node_ptr, v.seg_rel, (const float4*)T, att,           \
                                                           (const float4*)basis, (const float4*)root, x, geff, f_in, \
                                                           f_out, n_bases, datt_seg, d_x);                         \
    }
    bool tiled = false;
#define TILED_BWD(TFV, NBQV)                                                                                       \
    {                                                                                                              \
        auto kern = k_rgcn_node_bwd_tiled<TFV, NBQV>;                                                              \
        const size_t sm = node_bwd_tiled_smem<TFV, NBQV>(f_in);                                                    \
        if ((rc = ensure_dyn_smem((const void*)kern, sm))) return rc;                                              \
        kern<<<(unsigned)n_nodes, TILED_THREADS, sm, s>>>(v.node_ptr, v.seg_rel, (const float4*)T, (const float4*)att, \
                                                          (const float4*)basis, (const float4*)root, x, geff, f_in, \
                                                          datt_seg, d_x);                                          \
        tiled = true;                                                                                              \
    }
    static const bool tc_env = [] { const char* e = getenv("TIPB_RGCN_TC"); return !(e && (e[0] == '0' || e[0] == '1')); }();
    const bool use_tc = tc_env && n_rel >= RGCN_TC_MIN_REL;
#define TC_BWD(FOV, NBV)                                                                                           \
    {                                                                                                              \
        if ((rc = basis_y_launch(x, basis, (int)n_nodes, f_in, f_out, n_bases, Ybuf, s))) return rc;               \
        auto kern = k_rgcn_node_bwd_tc<FOV, NBV>;                                                                  \
        const size_t sm = RtBwdLayout<FOV, NBV>::BYTES;                                                            \
        if ((rc = ensure_dyn_smem((const void*)kern, sm))) return rc;                                              \
        kern<<<(unsigned)n_nodes, RT_THREADS, sm, s>>>(v.node_ptr, v.seg_rel, T, att, Ybuf, datt_seg, Qbuf,        \
                                                       rgcn_tc_error_flag(), rgcn_tc_dbg());                       \
        if ((rc = basis_dx_launch(Qbuf, basis, geff, root, (int)n_nodes, f_in, f_out, n_bases, dx_partial, d_x, s))) return rc; \
        tiled = true;                                                                                              \
    }
    if (use_tc && (f_in == 16 || f_in == 32 || f_in == 64 || f_in == 128)) {      // TIPB_RGCN_TC=1: forward pass only
        if (n_bases == 32) {
            if (f_out == 32) TC_BWD(32, 32) else if (f_out == 16) TC_BWD(16, 32)
        } else if (n_bases == 16) {
            if (f_out == 32) TC_BWD(32, 16) else if (f_out == 16) TC_BWD(16, 16)
        }
    }
#undef TC_BWD
    if (tiled) {
    } else if (n_bases == 32) {
        if (f_out == 64) TILED_BWD(16, 8) else if (f_out == 32) TILED_BWD(8, 8) else if (f_out == 16) TILED_BWD(4, 8)
    } else if (n_bases == 16) {
        if (f_out == 64) TILED_BWD(16, 4) else if (f_out == 32) TILED_BWD(8, 4) else if (f_out == 16) TILED_BWD(4, 4)
    }
#undef TILED_BWD
    if (!tiled) {
        switch (nb) {
            case 1: LAUNCH_BWD(1) break;
            case 2: LAUNCH_BWD(2) break;
            case 4: LAUNCH_BWD(4) break;
            default: LAUNCH_BWD(8) break;
        }
    }
#undef LAUNCH_BWD
    k_rel_reduce<<<(unsigned)n_rel, REL_REDUCE_THREADS, 0, s>>>(v.rel_seg_ptr, v.rel_seg, v.counts, datt_seg, n_bases, 1.0f, d_att);
    if ((rc = atb_launch(g_saved, ghat, (int)n_nodes, m_basis, f_out, d_basis, partial, s))) return rc;
    if ((rc = atb_launch(x, geff, (int)n_nodes, f_in, f_out, d_root, partial, s))) return rc;
    TIPB_CHECK_LAUNCH("rgcn_bwd");
    return TIPB_OK;
}
}
