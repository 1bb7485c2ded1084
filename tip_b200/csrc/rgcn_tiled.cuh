// Register-tiled node kernels of the R-GCN layer for the common shapes (feature widths that are
// multiples of 4 up to 64, n_bases in {16, 32}); rgcn.cu keeps generic kernels for everything else.
//
// Per node i the forward pass is the small product  G[i] = att_i^T [B x S_i] . H_i [S_i x F]  (S_i = number
// of non-empty (i, relation) segments, <= n_rel).  Each thread owns a 4(b) x 4(f) tile of G, so a segment
// costs two 128-bit shared-memory loads and eight packed FFMA2 per thread; the 256 threads of the CTA form
// NG = 256 / (F/4 * B/4) groups that split the segments (split-K) and are summed in a fixed order at the
// end.  Segment rows (H or T) and their relation's att rows are staged with cp.async, double buffered.
#pragma once
#include "common.cuh"

namespace tipb {

constexpr int TILED_THREADS = 256;
// One CTA per node: 645 CTAs on the drug graph.  At 4 resident CTAs per SM (592 slots) that is 1.09 waves -- the last
// 53 CTAs run alone; capping registers for 5 per SM (740 slots) keeps the whole grid resident in one wave.
#ifndef TILED_MIN_CTAS
#define TILED_MIN_CTAS 5
#endif
constexpr int CHT = 32;  // segments per staged chunk
// float4 slots of the staging area: two stages of payload + att rows, and never less than the
// TILED_THREADS * 4 slots the split-K partial tiles need when they reuse it
__host__ __device__ constexpr int tiled_stage_f4(int tp, int nbq) {
    return 2 * CHT * (tp + nbq) > TILED_THREADS * 4 ? 2 * CHT * (tp + nbq) : TILED_THREADS * 4;
}

__device__ __forceinline__ void cp_async16_t(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit_t() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_t() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// rows [s0, s0+cn) of `payload` (TP float4 each, contiguous) and the att rows of their relations (NBQ float4 each)
template <int TP, int NBQ>
__device__ __forceinline__ void stage_tiled(float4* sP, float4* sA, const float4* __restrict__ payload,
                                            const float4* __restrict__ att4, const int* __restrict__ seg_rel, int s0,
                                            int cn) {
    const float4* src = payload + int64_t(s0) * TP;
    for (int i = threadIdx.x; i < cn * TP; i += TILED_THREADS) cp_async16_t(&sP[i], &src[i]);
    for (int i = threadIdx.x; i < cn * NBQ; i += TILED_THREADS) {
        const int c = i / NBQ, q = i - c * NBQ;
        cp_async16_t(&sA[i], &att4[int64_t(seg_rel[s0 + c]) * NBQ + q]);
    }
    cp_async_commit_t();
}

// ------------------------------------------------------------------------------------------------ forward
// smem (float4 units): stage [2][CHT*TF] | att [2][CHT*NBQ] | Gs [B*TF] | then floats: xs [F_IN] | red [256]
// (the split-K partials alias the staging area once the main loop is done)
template <int TF, int NBQ>
__global__ void __launch_bounds__(TILED_THREADS, TILED_MIN_CTAS)
k_rgcn_node_fwd_tiled(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float* __restrict__ inv_deg,
                      const float4* __restrict__ H, const float4* __restrict__ att4, const float* __restrict__ basis,
                      const float* __restrict__ root, const float* __restrict__ bias, const float* __restrict__ x,
                      int f_out, int relu, float* __restrict__ out, float4* __restrict__ g_saved) {
    constexpr int F_IN = TF * 4, B = NBQ * 4, GT = TF * NBQ, NG = TILED_THREADS / GT;
    static_assert(GT <= TILED_THREADS && TILED_THREADS % GT == 0, "group size must divide the CTA");
    constexpr int STAGE = tiled_stage_f4(TF, NBQ);  // staging area, also holds the NG*B*TF split-K partials
    extern __shared__ float4 smem4[];
    float4* sH = smem4;
    float4* sA = sH + 2 * CHT * TF;
    float4* Gs4 = smem4 + STAGE;
    float* xs = reinterpret_cast<float*>(Gs4 + B * TF);
    float* red = xs + F_IN;
    const int tid = threadIdx.x;
    const int g = tid / GT, lt = tid % GT, tx = lt % TF, ty = lt / TF;

    const int i = blockIdx.x;
    const int sb = node_ptr[i], se = node_ptr[i + 1];
    const int n_chunks = (se - sb + CHT - 1) / CHT;
    for (int f = tid; f < F_IN; f += TILED_THREADS) xs[f] = x[int64_t(i) * F_IN + f];

    float4 acc0 = f4_zero(), acc1 = f4_zero(), acc2 = f4_zero(), acc3 = f4_zero();
    if (n_chunks > 0) stage_tiled<TF, NBQ>(sH, sA, H, att4, seg_rel, sb, min(CHT, se - sb));
    for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        if (c + 1 < n_chunks) {
            const int s1 = sb + (c + 1) * CHT;
            stage_tiled<TF, NBQ>(sH + (st ^ 1) * CHT * TF, sA + (st ^ 1) * CHT * NBQ, H, att4, seg_rel, s1,
                                 min(CHT, se - s1));
            cp_async_wait_t<1>();
        } else {
            cp_async_wait_t<0>();
        }
        __syncthreads();
        const int cn = min(CHT, se - (sb + c * CHT));
        const float4* h4 = sH + st * CHT * TF + tx;
        const float4* a4 = sA + st * CHT * NBQ + ty;
#pragma unroll 4
        for (int q = g; q < cn; q += NG) {
            const float4 h = h4[q * TF];
            const float4 a = a4[q * NBQ];
            acc0 = f4_fma(a.x, h, acc0);
            acc1 = f4_fma(a.y, h, acc1);
            acc2 = f4_fma(a.z, h, acc2);
            acc3 = f4_fma(a.w, h, acc3);
        }
        __syncthreads();
    }

    // split-K partials -> fixed-order sum -> G (shared + global)
    float4* part = smem4;  // [NG][B][TF]
    part[(g * B + ty * 4 + 0) * TF + tx] = acc0;
    part[(g * B + ty * 4 + 1) * TF + tx] = acc1;
    part[(g * B + ty * 4 + 2) * TF + tx] = acc2;
    part[(g * B + ty * 4 + 3) * TF + tx] = acc3;
    __syncthreads();
    for (int idx = tid; idx < B * TF; idx += TILED_THREADS) {
        float4 s = part[idx];
#pragma unroll
        for (int gg = 1; gg < NG; ++gg) s = f4_add(s, part[gg * B * TF + idx]);
        Gs4[idx] = s;
        g_saved[int64_t(i) * B * TF + idx] = s;
    }
    __syncthreads();

    // out_i[o] = inv_deg_i * sum_{b,f} G[b,f] basis[b,f,o] + sum_f x_i[f] root[f,o]
    const float* Gs = reinterpret_cast<const float*>(Gs4);
    const int o = tid % f_out, q = tid / f_out, nq = TILED_THREADS / f_out;
    float s0 = 0.f, s1 = 0.f;
    int bf = q;
    for (; bf + nq < B * F_IN; bf += 2 * nq) {
        s0 = fmaf(Gs[bf], basis[int64_t(bf) * f_out + o], s0);
        s1 = fmaf(Gs[bf + nq], basis[int64_t(bf + nq) * f_out + o], s1);
    }
    if (bf < B * F_IN) s0 = fmaf(Gs[bf], basis[int64_t(bf) * f_out + o], s0);
    float sum = (s0 + s1) * inv_deg[i];
    for (int f = q; f < F_IN; f += nq) sum = fmaf(xs[f], root[f * f_out + o], sum);
    red[tid] = sum;
    __syncthreads();
    if (tid < f_out) {
        float tot = 0.f;
        for (int qq = 0; qq < nq; ++qq) tot += red[qq * f_out + tid];
        if (bias) tot += bias[tid];
        if (relu) tot = fmaxf(tot, 0.f);
        out[int64_t(i) * f_out + tid] = tot;
    }
}

template <int TF, int NBQ>
static size_t node_fwd_tiled_smem() {
    return size_t(tiled_stage_f4(TF, NBQ) + NBQ * 4 * TF) * sizeof(float4) + size_t(TF * 4 + TILED_THREADS) * sizeof(float);
}

// ------------------------------------------------------------------------------------------------ backward
// four per-thread partial dots d0..d3 (for b = 4ty..4ty+3) summed over the TFO lanes that share (group, ty);
// on return the lane with tx == writer_tx(b_local) holds the total of b_local (see d_att_writer below)
template <int TFO>
__device__ __forceinline__ float reduce4_over_tx(float d0, float d1, float d2, float d3, int tx) {
    static_assert(TFO == 4 || TFO == 8 || TFO == 16, "unsupported payload width");
    // halving step 1: upper half of the lanes keeps (d2,d3), lower half keeps (d0,d1)
    const bool hi = tx & (TFO / 2);
    float a0 = hi ? d2 : d0, a1 = hi ? d3 : d1;
    const float b0 = hi ? d0 : d2, b1 = hi ? d1 : d3;
    a0 += __shfl_xor_sync(FULL, b0, TFO / 2);
    a1 += __shfl_xor_sync(FULL, b1, TFO / 2);
    // halving step 2
    const bool mid = tx & (TFO / 4);
    float c = mid ? a1 : a0;
    const float e = mid ? a0 : a1;
    c += __shfl_xor_sync(FULL, e, TFO / 4);
    // plain reduction over what is left of the tx bits
#pragma unroll
    for (int o = TFO / 8; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    return c;
}
// after reduce4_over_tx: lanes with (tx % (TFO/4)) == 0 hold b_local = 2*bit(TFO/2) + bit(TFO/4)
template <int TFO>
__device__ __forceinline__ bool d_att_writer(int tx, int& b_local) {
    b_local = ((tx & (TFO / 2)) ? 2 : 0) + ((tx & (TFO / 4)) ? 1 : 0);
    return (tx % (TFO / 4)) == 0;
}

// smem (float4 units): stage [2][CHT*TFO] | att [2][CHT*NBQ] | Ys [B*TFO] | Qs [B*TFO] | floats: Ds [CHT*B] | xs [f_in] | gs [F_OUT]
template <int TFO, int NBQ>
__global__ void __launch_bounds__(TILED_THREADS, TILED_MIN_CTAS)
k_rgcn_node_bwd_tiled(const int* __restrict__ node_ptr, const int* __restrict__ seg_rel, const float4* __restrict__ T,
                      const float4* __restrict__ att4, const float4* __restrict__ basis4,
                      const float4* __restrict__ root4, const float* __restrict__ x, const float* __restrict__ geff,
                      int f_in, float* __restrict__ datt_seg, float* __restrict__ d_x) {
    constexpr int F_OUT = TFO * 4, B = NBQ * 4, GT = TFO * NBQ, NG = TILED_THREADS / GT;
    static_assert(GT <= TILED_THREADS && TILED_THREADS % GT == 0, "group size must divide the CTA");
    constexpr int STAGE = tiled_stage_f4(TFO, NBQ);  // staging area, also holds the NG*B*TFO split-K partials
    extern __shared__ float4 smem4[];
    float4* sT = smem4;
    float4* sA = sT + 2 * CHT * TFO;
    float4* Ys4 = smem4 + STAGE;
    float4* Qs4 = Ys4 + B * TFO;
    float* Ds = reinterpret_cast<float*>(Qs4 + B * TFO);
    float* xs = Ds + CHT * B;
    float* gs = xs + f_in;  // f_in % 4 == 0 keeps 16-byte alignment
    const int tid = threadIdx.x;
    const int g = tid / GT, lt = tid % GT, tx = lt % TFO, ty = lt / TFO;

    const int j = blockIdx.x;
    const int sb = node_ptr[j], se = node_ptr[j + 1];
    const int n_chunks = (se - sb + CHT - 1) / CHT;
    for (int f = tid; f < f_in; f += TILED_THREADS) xs[f] = x[int64_t(j) * f_in + f];
    for (int o = tid; o < F_OUT; o += TILED_THREADS) gs[o] = geff[int64_t(j) * F_OUT + o];
    if (n_chunks > 0) stage_tiled<TFO, NBQ>(sT, sA, T, att4, seg_rel, sb, min(CHT, se - sb));
    __syncthreads();

    // Y[j,b,o4] = sum_f x_j[f] basis[b,f,o4], computed once per CTA
    if (n_chunks > 0) {
        for (int idx = tid; idx < B * TFO; idx += TILED_THREADS) {
            const int b = idx / TFO, o4 = idx - b * TFO;
            const float4* bp = basis4 + (int64_t(b) * f_in) * TFO + o4;
            float4 y0 = f4_zero(), y1 = f4_zero();
            int f = 0;
            for (; f + 1 < f_in; f += 2) {
                y0 = f4_fma(xs[f], bp[int64_t(f) * TFO], y0);
                y1 = f4_fma(xs[f + 1], bp[int64_t(f + 1) * TFO], y1);
            }
            if (f < f_in) y0 = f4_fma(xs[f], bp[int64_t(f) * TFO], y0);
            Ys4[idx] = f4_add(y0, y1);
        }
    }
    __syncthreads();
    const float4 y0 = Ys4[(ty * 4 + 0) * TFO + tx], y1 = Ys4[(ty * 4 + 1) * TFO + tx];
    const float4 y2 = Ys4[(ty * 4 + 2) * TFO + tx], y3 = Ys4[(ty * 4 + 3) * TFO + tx];
    int b_local;
    const bool writer = d_att_writer<TFO>(tx, b_local);

    float4 q0 = f4_zero(), q1 = f4_zero(), q2 = f4_zero(), q3 = f4_zero();
    for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        const int s0 = sb + c * CHT;
        if (c + 1 < n_chunks) {
            const int s1 = s0 + CHT;
            stage_tiled<TFO, NBQ>(sT + (st ^ 1) * CHT * TFO, sA + (st ^ 1) * CHT * NBQ, T, att4, seg_rel, s1,
                                  min(CHT, se - s1));
            cp_async_wait_t<1>();
        } else {
            cp_async_wait_t<0>();
        }
        __syncthreads();
        const int cn = min(CHT, se - s0);
        const float4* t4 = sT + st * CHT * TFO + tx;
        const float4* a4 = sA + st * CHT * NBQ + ty;
        // every group runs the same number of trips so that the shuffles stay warp-uniform
        for (int q = g; q < CHT; q += NG) {
            const bool live = q < cn;
            const float4 t = live ? t4[q * TFO] : f4_zero();
            const float4 a = live ? a4[q * NBQ] : f4_zero();
            q0 = f4_fma(a.x, t, q0);
            q1 = f4_fma(a.y, t, q1);
            q2 = f4_fma(a.z, t, q2);
            q3 = f4_fma(a.w, t, q3);
            const float d = reduce4_over_tx<TFO>(f4_dot(t, y0), f4_dot(t, y1), f4_dot(t, y2), f4_dot(t, y3), tx);
            if (live && writer) Ds[q * B + ty * 4 + b_local] = d;
        }
        __syncthreads();
        for (int idx = tid; idx < cn * B; idx += TILED_THREADS) datt_seg[int64_t(s0) * B + idx] = Ds[idx];
        // the barrier at the top of the next trip orders these reads before Ds is rewritten
    }
    __syncthreads();

    float4* part = smem4;  // [NG][B][TFO]
    part[(g * B + ty * 4 + 0) * TFO + tx] = q0;
    part[(g * B + ty * 4 + 1) * TFO + tx] = q1;
    part[(g * B + ty * 4 + 2) * TFO + tx] = q2;
    part[(g * B + ty * 4 + 3) * TFO + tx] = q3;
    __syncthreads();
    for (int idx = tid; idx < B * TFO; idx += TILED_THREADS) {
        float4 s = part[idx];
#pragma unroll
        for (int gg = 1; gg < NG; ++gg) s = f4_add(s, part[gg * B * TFO + idx]);
        Qs4[idx] = s;
    }
    __syncthreads();

    // dX_j[f] = sum_{b,o} Q[b,o] basis[b,f,o] + sum_o geff_j[o] root[f,o]
    const int o4 = tid % TFO, fsel = tid / TFO;
    const float4 g4 = reinterpret_cast<const float4*>(gs)[o4];
    for (int f = fsel; f < f_in; f += TILED_THREADS / TFO) {
        float s0 = f4_dot(g4, root4[f * TFO + o4]), s1 = 0.f;
#pragma unroll 4
        for (int b = 0; b < B; b += 2) {
            s0 += f4_dot(Qs4[b * TFO + o4], basis4[(int64_t(b) * f_in + f) * TFO + o4]);
            s1 += f4_dot(Qs4[(b + 1) * TFO + o4], basis4[(int64_t(b + 1) * f_in + f) * TFO + o4]);
        }
        float sum = s0 + s1;
#pragma unroll
        for (int off = TFO >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);
        if (o4 == 0) d_x[int64_t(j) * f_in + f] = sum;
    }
}

template <int TFO, int NBQ>
static size_t node_bwd_tiled_smem(int f_in) {
    return size_t(tiled_stage_f4(TFO, NBQ) + 2 * NBQ * 4 * TFO) * sizeof(float4) +
           size_t(CHT * NBQ * 4 + f_in + TFO * 4) * sizeof(float);
}

}  // namespace tipb
