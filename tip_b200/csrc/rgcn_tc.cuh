// R-GCN node contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Per node i the forward pass of MyRGCNConv2 (reference src/layers.py:157-188, after the segment sums H) is the product
//     G[i] = att_i^T [B x S_i] . H_i [S_i x F]          S_i = non-empty (i, relation) segments of the node, <= n_rel
// which round 1 ran on CUDA cores (rgcn_tiled.cuh: 17.8 TFLOP/s, 24 % of the fp32 FMA peak -- the four node launches were
// 19 % of the step).  Here it is one accumulation chain of tcgen05.mma per node:
//     D[f, b] (TMEM, M = 64 rows f, N = B columns)  +=  A[f, k] * Bm[b, k],   k = segment (K = 16 per instruction)
// A = H_i^T and Bm = att rows of the node's segments, both converted to three bf16 pieces per fp32 value while they are
// staged (x = hi + mid + lo: 24 significant bits) in the K-major no-swizzle canonical layout; the six piece products
// hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi select their pieces through the descriptor start addresses, so the result
// carries fp32 accuracy (dropped terms <= 2^-24 relative).  Chunks of 64 segments are double buffered: all eight warps
// convert chunk c+1 (coalesced loads: lane <-> feature column, eight consecutive segments per thread -> one 16-byte store
// per piece) while the tensor core works on chunk c; tcgen05.commit on the stage's mbarrier frees it.  One CTA per node,
// two CTAs per SM (each allocates 32 TMEM columns).  M = 64 uses the half-subpartition TMEM layout: row f lives in lane
// (f % 16) + 32 (f / 16), so warp w < 4 drains rows 16 w .. 16 w + 15 with one tcgen05.ld.
// The kernel writes G (g_saved); out_i = inv_deg_i <G_i, basis> + x_i root is a GEMM batched over the nodes
// (rgcn_dense.cuh) -- inside the node's CTA it would re-read the whole basis tensor per node.
// Every barrier wait is bounded: a protocol fault sets a flag (tipb_rgcn_tc_status) instead of hanging the GPU.
#pragma once
#include "umma.cuh"

namespace tipb {

constexpr int RT_THREADS = 256;
constexpr int RT_KC = 64;                       // segments per chunk = 4 MMA K-steps
constexpr int RT_SBO = (RT_KC / 8) * 128;       // bytes between 8-row groups: 8 chunks of 16 bytes x 8 rows
constexpr int RT_M = 64;                        // MMA M (feature rows, zero rows above F)

__device__ int g_rgcn_tc_error = 0;

__device__ __forceinline__ uint32_t rt_chunk_offset(int row, int chunk8) {
    return uint32_t((row >> 3) * RT_SBO + chunk8 * 128 + (row & 7) * 16);
}

// smem: [2 stages] x { A pieces 3 x (8 row groups x SBO) | B pieces 3 x (NB/8 row groups x SBO) }
template <int F, int NB>
struct RtLayout {
    static constexpr int A_PIECE = (RT_M / 8) * RT_SBO;          // 8 KB
    static constexpr int B_PIECE = (NB / 8) * RT_SBO;            // 4 KB (NB = 32)
    static constexpr int STAGE = 3 * A_PIECE + 3 * B_PIECE;
    static constexpr int NS = 2;       // operand stages.  Measured with 3 (216 KB of the SM's 256 KB L1/shared array): 67 -> 79 us
    static constexpr size_t BYTES = NS * size_t(STAGE) + 1024;
};

template <int F, int NB>
__global__ void __launch_bounds__(RT_THREADS, 2)
k_rgcn_node_fwd_tc(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float* __restrict__ H,
                   const float* __restrict__ att, float* __restrict__ g_saved, int* __restrict__ error_flag, int dbg) {
    // dbg (measurements only, TIPB_RGCN_TC_DBG): bits 0-2 = piece products to issue (0 = all six), bit 3 = no MMAs,
    // bit 4 = no conversion / staging stores, bit 5 = no operand loads
    using LY = RtLayout<F, NB>;
    static_assert(F == 64 || F == 32 || F == 16, "feature width");
    static_assert(NB == 32 || NB == 16, "number of bases");
    constexpr int A_UNITS = F * (RT_KC / 8) / RT_THREADS;        // (feature, 8-segment octet) units per thread: 2 / 1 / 0.5
    constexpr int A_ITERS = A_UNITS > 0 ? A_UNITS : 1;
    extern __shared__ __align__(1024) uint8_t rt_smem[];
    uint8_t* stage0 = rt_smem;
    __shared__ uint64_t bar_free[LY::NS], bar_done;
    __shared__ uint32_t s_tmem_base;

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int i = blockIdx.x;
    const int sb = node_ptr[i], se = node_ptr[i + 1];
    const int n_chunks = (se - sb + RT_KC - 1) / RT_KC;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < LY::NS; ++q) mbar_init(&bar_free[q], 1);
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (wid == 0 && !(dbg & 64)) tmem_alloc(&s_tmem_base, 32);
    // rows f >= F of the A tiles are never written and must read as zero: row groups F / 8 .. 7 are one contiguous byte
    // range of every piece (nothing to do for F = 64); the measurement modes zero everything
    if (dbg) {
        if (!(dbg & 128))
        for (int k = tid; k < LY::NS * LY::STAGE / 16; k += RT_THREADS) reinterpret_cast<uint4*>(stage0)[k] = make_uint4(0u, 0u, 0u, 0u);
    } else if (F < RT_M) {
        constexpr int Z16 = F < RT_M ? (RT_M - F) / 8 * RT_SBO / 16 : 1;   // 16-byte words per piece (branch is dead for F = 64)
        for (int k = tid; k < 3 * LY::NS * Z16; k += RT_THREADS) {
            const int piece = k / Z16, w = k % Z16;               // piece = stage * 3 + piece index
            uint8_t* base = stage0 + (piece / 3) * LY::STAGE + (piece % 3) * LY::A_PIECE + (F / 8) * RT_SBO;
            reinterpret_cast<uint4*>(base)[w] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // ---- raw values of one chunk in registers: A units (feature af, octet ao + 4 u), one B unit (base bb, octet bo)
    const int af = tid % F, ao = tid / F;               // F = 64: ao in 0..3 (+4 for the second unit); F = 32: 0..7; F = 16: 0..15
    const int bb = tid % NB, bo = tid / NB;             // NB = 32: bo in 0..7;  NB = 16: 0..15 (only bo < 8 is used)
    const bool a_live = ((F * (RT_KC / 8) >= RT_THREADS) || ao < RT_KC / 8) && !(dbg & 32);
    const bool b_live = bo < RT_KC / 8 && !(dbg & 32);
    const int n_prod = (dbg & 7) ? (dbg & 7) : 6;
    // two register sets: the loads of chunk c + 2 are issued right after chunk c is converted, so every load has a whole
    // chunk trip (conversion, barrier, MMA hand-off) to land before its values are needed
    float ra0[A_ITERS][8], rb0[8], ra1[A_ITERS][8], rb1[8];
    // Full chunks (all but a node's last) take unpredicated loads at compile-time offsets from two running pointers; the
    // first version spent ~7 instructions per load on per-element bounds predicates and 64-bit index arithmetic and was
    // issue bound at two CTAs per SM.
    const float* const pa0 = H + int64_t(sb + ao * 8) * F + af;
    const int* const pr0 = seg_rel + sb + bo * 8;
    auto load_raw = [&](int c, float (&ra)[A_ITERS][8], float (&rb)[8]) {
        const int s0 = sb + c * RT_KC;
        const float* pa = pa0 + int64_t(c) * (RT_KC * F);
        const int* pr = pr0 + c * RT_KC;
        if (s0 + RT_KC <= se) {
            if (a_live) {
#pragma unroll
                for (int u = 0; u < A_ITERS; ++u)
#pragma unroll
                    for (int j = 0; j < 8; ++j) ra[u][j] = pa[(u * (RT_THREADS / F) * 8 + j) * F];
            }
            if (b_live) {
                int rel[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) rel[j] = pr[j];
#pragma unroll
                for (int j = 0; j < 8; ++j) rb[j] = att[rel[j] * NB + bb];
            }
        } else {
#pragma unroll
            for (int u = 0; u < A_ITERS; ++u) {
                const int oct = ao + u * (RT_THREADS / F);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int s = s0 + oct * 8 + j;
                    ra[u][j] = (a_live && s < se) ? pa[(u * (RT_THREADS / F) * 8 + j) * F] : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = s0 + bo * 8 + j;
                rb[j] = (b_live && s < se) ? att[pr[j] * NB + bb] : 0.f;
            }
        }
    };
    auto store_pieces = [&](uint8_t* st, const float (&ra)[A_ITERS][8], const float (&rb)[8]) {
#pragma unroll
        for (int u = 0; u < A_ITERS; ++u) {
            if (!a_live) continue;
            const int oct = ao + u * (RT_THREADS / F);
            const Pieces8 p = split8(ra[u]);
            const uint32_t off = rt_chunk_offset(af, oct);
            *reinterpret_cast<uint4*>(st + off) = p.hi;
            *reinterpret_cast<uint4*>(st + LY::A_PIECE + off) = p.mid;
            *reinterpret_cast<uint4*>(st + 2 * LY::A_PIECE + off) = p.lo;
        }
        if (b_live) {
            const Pieces8 p = split8(rb);
            uint8_t* sbm = st + 3 * LY::A_PIECE;
            const uint32_t off = rt_chunk_offset(bb, bo);
            *reinterpret_cast<uint4*>(sbm + off) = p.hi;
            *reinterpret_cast<uint4*>(sbm + LY::B_PIECE + off) = p.mid;
            *reinterpret_cast<uint4*>(sbm + 2 * LY::B_PIECE + off) = p.lo;
        }
    };

    bool ok = true;
    if (n_chunks > 0) load_raw(0, ra0, rb0);
    if (n_chunks > 1) load_raw(1, ra1, rb1);
    constexpr uint32_t idesc = umma_idesc_bf16(RT_M, NB);
    // products hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi (piece index 0 = hi, 1 = mid, 2 = lo)
    constexpr int piece_a[6] = {0, 0, 1, 1, 0, 2};
    constexpr int piece_b[6] = {0, 1, 0, 1, 2, 0};
    auto chunk_trip = [&](int c, float (&ra)[A_ITERS][8], float (&rb)[8]) {
        // the MMAs of chunk c - NS must have read this stage (a third stage did not pay: see RtLayout::NS)
        const int st = c % LY::NS;
        uint8_t* stage = stage0 + st * LY::STAGE;
        if (c >= LY::NS) ok = mbar_wait(&bar_free[st], uint32_t(c / LY::NS - 1) & 1u);
        if (!ok) return;
        if (!(dbg & 16)) store_pieces(stage, ra, rb);
        if (c + 2 < n_chunks) load_raw(c + 2, ra, rb);
        fence_proxy_async();                             // generic-proxy stores -> visible to the async proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(stage), b_base = a_base + 3 * LY::A_PIECE;
            if (!(dbg & 8))
#pragma unroll
            for (int kk = 0; kk < RT_KC / 16; ++kk) {
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const uint64_t da = umma_desc(a_base + piece_a[p] * LY::A_PIECE + kk * 256, 128, RT_SBO);
                    const uint64_t db = umma_desc(b_base + piece_b[p] * LY::B_PIECE + kk * 256, 128, RT_SBO);
                    if (p < n_prod) umma_bf16(tmem_base, da, db, idesc, (c > 0 || kk > 0 || p > 0) ? 1u : 0u);
                }
            }
            umma_commit(&bar_free[st]);
            if (c + 1 == n_chunks) umma_commit(&bar_done);
        }
    };
    for (int c = 0; c < n_chunks && ok; c += 2) {
        chunk_trip(c, ra0, rb0);
        if (ok && c + 1 < n_chunks) chunk_trip(c + 1, ra1, rb1);
    }
    if (ok && n_chunks > 0) ok = mbar_wait(&bar_done, 0);
    if (!ok) atomicExch(error_flag, 1);
    tc_fence_after();

    // ---- G: TMEM -> global.  Row f lives in lane (f % 16) + 32 (f / 16): warp w < F / 16 holds rows 16 w .. + 15
    if (wid < F / 16) {
        uint32_t v[32];
        if (n_chunks > 0 && ok && !(dbg & 64)) {
            tmem_ld32(tmem_base + (uint32_t(wid * 32) << 16), v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
        if (lane < 16) {
            const int f = wid * 16 + lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) g_saved[(int64_t(i) * NB + b) * F + f] = __uint_as_float(v[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0 && !(dbg & 64)) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

// ================================================================================================ backward
// Per node j (by-source plan): T[s] = segment sums of the scaled output gradient (S_j x FO).
//   Q[j,b,:]      = sum_s att[r_s,b] T[s]              D1[o, b] += A1[o, k = s] * B1[b, k = s]     (as the forward pass)
//   datt_seg[s,b] = <T[s], Y_j[b]>,  Y_j = x_j basis   D2[s, b]  = A2[s, k = o] * B2[b, k = o]     (per chunk, not accumulated)
//   Y before and dX_j = <Q_j, basis^T> + geff_j root^T after are batched over the nodes (rgcn_dense.cuh)
// D2 of chunk c is drained from TMEM (two alternating 32-column accumulators) while the tensor core works on chunk c + 1.
template <int FO, int NB>
struct RtBwdLayout {
    static constexpr int A1_PIECE = (RT_M / 8) * RT_SBO;         // rows o (64, zero above FO), K = 64 segments
    static constexpr int B1_PIECE = (NB / 8) * RT_SBO;           // rows b, K = 64 segments
    static constexpr int SBO2 = (FO / 8) * 128;                  // K = FO features: FO / 8 chunks per row
    static constexpr int A2_PIECE = (RT_KC / 8) * SBO2;          // rows s (64), K = FO
    static constexpr int B2_PIECE = (NB / 8) * SBO2;             // rows b, K = FO  (Y_j, built once per node)
    static constexpr int STAGE = 3 * (A1_PIECE + B1_PIECE + A2_PIECE);
    static constexpr size_t BYTES = 2 * size_t(STAGE) + 3 * size_t(B2_PIECE) + 1024;
};

template <int FO, int NB>
__global__ void __launch_bounds__(RT_THREADS, 2)
k_rgcn_node_bwd_tc(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float* __restrict__ T,
                   const float* __restrict__ att, const float* __restrict__ Y, float* __restrict__ datt_seg,
                   float* __restrict__ Q, int* __restrict__ error_flag, int dbg) {
    using LY = RtBwdLayout<FO, NB>;
    static_assert(FO == 32 || FO == 16, "payload width");
    static_assert(NB == 32 || NB == 16, "number of bases");
    extern __shared__ __align__(1024) uint8_t rt_smem[];
    uint8_t* stage0 = rt_smem;
    uint8_t* sB2 = rt_smem + 2 * LY::STAGE;
    __shared__ uint64_t bar_free[2];
    __shared__ uint32_t s_tmem_base;

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x;
    const int sb = node_ptr[j], se = node_ptr[j + 1];
    const int n_chunks = (se - sb + RT_KC - 1) / RT_KC;

    if (tid == 0) {
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
        fence_barrier_init();
    }
    if (wid == 0) tmem_alloc(&s_tmem_base, 128);
    // rows o >= FO of the A1 tiles are never written and must read as zero (one contiguous byte range per piece); every
    // other operand row is rewritten by each chunk.  The measurement modes zero everything.
    if (dbg) {
        for (int k = tid; k < 2 * LY::STAGE / 16; k += RT_THREADS) reinterpret_cast<uint4*>(stage0)[k] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        constexpr int Z16 = (RT_M - FO) / 8 * RT_SBO / 16;
        for (int k = tid; k < 6 * Z16; k += RT_THREADS) {
            const int piece = k / Z16, w = k % Z16;
            uint8_t* base = stage0 + (piece / 3) * LY::STAGE + (piece % 3) * LY::A1_PIECE + (FO / 8) * RT_SBO;
            reinterpret_cast<uint4*>(base)[w] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // B2 pieces (rows b, K = o) from Y[j] = x_j basis (rgcn_dense.cuh: k_basis_y)
    if (n_chunks > 0 && tid < NB * (FO / 8)) {
        const int b = tid / (FO / 8), oc = tid % (FO / 8);
        const float4* y4 = reinterpret_cast<const float4*>(Y + (int64_t(j) * NB + b) * FO + oc * 8);
        const float4 u0 = y4[0], u1 = y4[1];
        const float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        const Pieces8 p = split8(v);
        const uint32_t off = uint32_t((b >> 3) * LY::SBO2 + oc * 128 + (b & 7) * 16);
        *reinterpret_cast<uint4*>(sB2 + off) = p.hi;
        *reinterpret_cast<uint4*>(sB2 + LY::B2_PIECE + off) = p.mid;
        *reinterpret_cast<uint4*>(sB2 + 2 * LY::B2_PIECE + off) = p.lo;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // ---- raw values of one chunk: A1 unit (column ao1, octet a1o), B1 unit (base bb, octet bo), A2 unit (segment as, octet a2o)
    const int ao1 = tid % FO, a1o = tid / FO;              // FO = 32: octets 0..7; FO = 16: 0..15 (only < 8 used)
    const int bb = tid % NB, bo = tid / NB;
    const int a2s = tid / (FO / 8), a2o = tid % (FO / 8);  // FO = 32: segments 0..63; FO = 16: 0..127 (only < 64 used)
    const bool a1_live = a1o < RT_KC / 8 && !(dbg & 32), b_live = bo < RT_KC / 8 && !(dbg & 32), a2_live = a2s < RT_KC && !(dbg & 32);
    const int n_prod = (dbg & 7) ? (dbg & 7) : 6;
    float ra0[8], rb0[8], r20[8], ra1[8], rb1[8], r21[8];     // two register sets, as in the forward kernel
    // full chunks: unpredicated loads at compile-time offsets from running pointers (see the forward kernel)
    const float* const pa0 = T + int64_t(sb + a1o * 8) * FO + ao1;
    const int* const pr0 = seg_rel + sb + bo * 8;
    const float* const p20 = T + int64_t(sb + a2s) * FO + a2o * 8;
    auto load_raw = [&](int c, float (&ra)[8], float (&rb)[8], float (&r2)[8]) {
        const int s0 = sb + c * RT_KC;
        const float* pa = pa0 + int64_t(c) * (RT_KC * FO);
        const int* pr = pr0 + c * RT_KC;
        const float4* t4 = reinterpret_cast<const float4*>(p20 + int64_t(c) * (RT_KC * FO));
        if (s0 + RT_KC <= se) {
            if (a1_live) {
#pragma unroll
                for (int q = 0; q < 8; ++q) ra[q] = pa[q * FO];
            }
            if (b_live) {
                int rel[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) rel[q] = pr[q];
#pragma unroll
                for (int q = 0; q < 8; ++q) rb[q] = att[rel[q] * NB + bb];
            }
            if (a2_live) {
                const float4 u0 = t4[0], u1 = t4[1];
                r2[0] = u0.x; r2[1] = u0.y; r2[2] = u0.z; r2[3] = u0.w; r2[4] = u1.x; r2[5] = u1.y; r2[6] = u1.z; r2[7] = u1.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int s1 = s0 + a1o * 8 + q, s2 = s0 + bo * 8 + q;
                ra[q] = (a1_live && s1 < se) ? pa[q * FO] : 0.f;
                rb[q] = (b_live && s2 < se) ? att[pr[q] * NB + bb] : 0.f;
            }
            if (a2_live && s0 + a2s < se) {
                const float4 u0 = t4[0], u1 = t4[1];
                r2[0] = u0.x; r2[1] = u0.y; r2[2] = u0.z; r2[3] = u0.w; r2[4] = u1.x; r2[5] = u1.y; r2[6] = u1.z; r2[7] = u1.w;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) r2[q] = 0.f;
            }
        }
    };
    auto store_pieces = [&](uint8_t* st, const float (&ra)[8], const float (&rb)[8], const float (&r2)[8]) {
        uint8_t* sA1 = st;
        uint8_t* sB1 = st + 3 * LY::A1_PIECE;
        uint8_t* sA2 = sB1 + 3 * LY::B1_PIECE;
        if (a1_live) {
            const Pieces8 p = split8(ra);
            const uint32_t off = rt_chunk_offset(ao1, a1o);
            *reinterpret_cast<uint4*>(sA1 + off) = p.hi;
            *reinterpret_cast<uint4*>(sA1 + LY::A1_PIECE + off) = p.mid;
            *reinterpret_cast<uint4*>(sA1 + 2 * LY::A1_PIECE + off) = p.lo;
        }
        if (b_live) {
            const Pieces8 p = split8(rb);
            const uint32_t off = rt_chunk_offset(bb, bo);
            *reinterpret_cast<uint4*>(sB1 + off) = p.hi;
            *reinterpret_cast<uint4*>(sB1 + LY::B1_PIECE + off) = p.mid;
            *reinterpret_cast<uint4*>(sB1 + 2 * LY::B1_PIECE + off) = p.lo;
        }
        if (a2_live) {
            const Pieces8 p = split8(r2);
            const uint32_t off = uint32_t((a2s >> 3) * LY::SBO2 + a2o * 128 + (a2s & 7) * 16);
            *reinterpret_cast<uint4*>(sA2 + off) = p.hi;
            *reinterpret_cast<uint4*>(sA2 + LY::A2_PIECE + off) = p.mid;
            *reinterpret_cast<uint4*>(sA2 + 2 * LY::A2_PIECE + off) = p.lo;
        }
    };
    // D2 of chunk c: rows s = 16 w + lane (lane < 16) of warp w < 4 -> datt_seg[s0 + s, :]
    auto drain_d2 = [&](int c) -> bool {
        if (wid >= 4) return true;
        if (!mbar_wait(&bar_free[c & 1], (c >> 1) & 1u)) return false;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(wid * 32) << 16) + 32 + (c & 1) * 32, v);
        tmem_ld_wait();
        const int s = sb + c * RT_KC + wid * 16 + lane;
        if (lane < 16 && s < se) {
            float4* dst = reinterpret_cast<float4*>(datt_seg + int64_t(s) * NB);
#pragma unroll
            for (int b4 = 0; b4 < NB / 4; ++b4)
                dst[b4] = make_float4(__uint_as_float(v[4 * b4]), __uint_as_float(v[4 * b4 + 1]), __uint_as_float(v[4 * b4 + 2]),
                                      __uint_as_float(v[4 * b4 + 3]));
        }
        tc_fence_before();
        return true;
    };

    bool ok = true;
    if (n_chunks > 0) load_raw(0, ra0, rb0, r20);
    if (n_chunks > 1) load_raw(1, ra1, rb1, r21);
    constexpr uint32_t idesc = umma_idesc_bf16(RT_M, NB);
    constexpr int piece_a[6] = {0, 0, 1, 1, 0, 2};
    constexpr int piece_b[6] = {0, 1, 0, 1, 2, 0};
    auto chunk_trip = [&](int c, float (&ra)[8], float (&rb)[8], float (&r2)[8]) {
        const int st = c & 1;
        uint8_t* stage = stage0 + st * LY::STAGE;
        if (c >= 2) ok = mbar_wait(&bar_free[st], ((c >> 1) - 1) & 1u);
        if (!ok) return;
        if (!(dbg & 16)) store_pieces(stage, ra, rb, r2);
        if (c + 2 < n_chunks) load_raw(c + 2, ra, rb, r2);
        fence_proxy_async();
        __syncthreads();                                   // (also: D2 of chunk c - 2 was drained in the previous trip)
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a1 = smem_u32(stage), b1 = a1 + 3 * LY::A1_PIECE, a2 = b1 + 3 * LY::B1_PIECE, b2 = smem_u32(sB2);
            if (!(dbg & 8))
#pragma unroll
            for (int kk = 0; kk < RT_KC / 16; ++kk) {
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const uint64_t da = umma_desc(a1 + piece_a[p] * LY::A1_PIECE + kk * 256, 128, RT_SBO);
                    const uint64_t db = umma_desc(b1 + piece_b[p] * LY::B1_PIECE + kk * 256, 128, RT_SBO);
                    if (p < n_prod) umma_bf16(tmem_base, da, db, idesc, (c > 0 || kk > 0 || p > 0) ? 1u : 0u);
                }
            }
            if (!(dbg & 8))
#pragma unroll
            for (int kk = 0; kk < FO / 16; ++kk) {
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const uint64_t da = umma_desc(a2 + piece_a[p] * LY::A2_PIECE + kk * 256, 128, LY::SBO2);
                    const uint64_t db = umma_desc(b2 + piece_b[p] * LY::B2_PIECE + kk * 256, 128, LY::SBO2);
                    if (p < n_prod) umma_bf16(tmem_base + 32 + st * 32, da, db, idesc, (kk > 0 || p > 0) ? 1u : 0u);
                }
            }
            umma_commit(&bar_free[st]);
        }
        if (c >= 1) ok = drain_d2(c - 1);                  // while the tensor core works on chunk c
    };
    for (int c = 0; c < n_chunks && ok; c += 2) {
        chunk_trip(c, ra0, rb0, r20);
        if (ok && c + 1 < n_chunks) chunk_trip(c + 1, ra1, rb1, r21);
    }
    if (ok && n_chunks > 0) ok = drain_d2(n_chunks - 1);   // its completion = every MMA of the node is done
    if (ok && n_chunks > 0 && wid >= 4) ok = mbar_wait(&bar_free[(n_chunks - 1) & 1], ((n_chunks - 1) >> 1) & 1u);
    if (!ok) atomicExch(error_flag, 2);
    tc_fence_after();

    // ---- Q: TMEM -> global (dX_j = <Q_j, basis^T> + geff_j root^T is batched over the nodes: rgcn_dense.cuh).
    // Row o lives in lane (o % 16) + 32 (o / 16)
    if (wid < FO / 16) {
        uint32_t v[32];
        if (n_chunks > 0 && ok) {
            tmem_ld32(tmem_base + (uint32_t(wid * 32) << 16), v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
        if (lane < 16) {
            const int o = wid * 16 + lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) Q[(int64_t(j) * NB + b) * FO + o] = __uint_as_float(v[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace tipb
