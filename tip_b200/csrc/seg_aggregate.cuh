// Segmented gather-reduce: the "relation-grouped segmented SpMM" of north_star item (3).
//
//   out[s, :] = sum_{p in [seg_ptr[s], seg_ptr[s+1])} scale[other[p]] * feat[other[p], :]
//
// One warp per (node, relation) segment.  A feature row of F floats is read by LPR = F/4 lanes as
// 128-bit words, so a warp keeps 32/LPR edges in flight per instruction.  When the whole feature
// matrix fits in shared memory (645 x 64 fp32 = 165 KB on the drug graph -- B200 allows 227 KB per
// CTA) it is staged there once per CTA and every gather is an LDS.128; otherwise rows come from
// L2 through the read-only path.  DRAM traffic is then just the index stream (4 B per edge).
// Edge indices are fetched 32 at a time (coalesced) and the next segment's bounds and first index
// block are prefetched while the current one is being reduced.
#pragma once
#include "common.cuh"

namespace tipb {

// Sum of (scaled) feature rows over the entries [beg,end) of one segment, computed by a whole warp.
// `idx` holds this lane's prefetched index other[beg+lane] (0 past the end).  Returns the full sum
// on every lane of every group (lane l of each group owns float4 l of the row).
template <int LPR, bool STAGED>
__device__ __forceinline__ float4 warp_gather_sum(const float4* __restrict__ src, const int* __restrict__ other,
                                                  const float* __restrict__ row_scale, int beg, int end, int idx) {
    constexpr int G = 32 / LPR;
    const int lane = lane_id();
    const int g = lane / LPR, l = lane % LPR;
    const bool scaled = !STAGED && row_scale != nullptr;
    float4 a0 = f4_zero(), a1 = f4_zero(), a2 = f4_zero(), a3 = f4_zero();
    for (int base = beg; base < end; base += 32) {
        if (base != beg) idx = (base + lane < end) ? ld_stream_i32(other + base + lane) : 0;
        const int cnt = min(32, end - base);
        int k0 = 0;
        // 4 x G entries per trip, four independent accumulators
        for (; k0 + 4 * G <= cnt; k0 += 4 * G) {
            int j0 = __shfl_sync(FULL, idx, k0 + g);
            int j1 = __shfl_sync(FULL, idx, k0 + G + g);
            int j2 = __shfl_sync(FULL, idx, k0 + 2 * G + g);
            int j3 = __shfl_sync(FULL, idx, k0 + 3 * G + g);
            float4 v0 = src[j0 * LPR + l], v1 = src[j1 * LPR + l], v2 = src[j2 * LPR + l], v3 = src[j3 * LPR + l];
            if (scaled) {
                a0 = f4_fma(row_scale[j0], v0, a0); a1 = f4_fma(row_scale[j1], v1, a1);
                a2 = f4_fma(row_scale[j2], v2, a2); a3 = f4_fma(row_scale[j3], v3, a3);
            } else {
                a0 = f4_add(a0, v0); a1 = f4_add(a1, v1); a2 = f4_add(a2, v2); a3 = f4_add(a3, v3);
            }
        }
        for (; k0 < cnt; k0 += G) {
            const int k = k0 + g;
            int j = __shfl_sync(FULL, idx, k & 31);
            if (k < cnt) {
                float4 v = src[j * LPR + l];
                if (scaled) a0 = f4_fma(row_scale[j], v, a0);
                else a0 = f4_add(a0, v);
            }
        }
    }
    float4 a = f4_add(f4_add(a0, a1), f4_add(a2, a3));
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        a.x += __shfl_xor_sync(FULL, a.x, o);
        a.y += __shfl_xor_sync(FULL, a.y, o);
        a.z += __shfl_xor_sync(FULL, a.z, o);
        a.w += __shfl_xor_sync(FULL, a.w, o);
    }
    return a;
}

// Staged variant: the feature matrix sits in shared memory followed by ONE ALL-ZERO ROW (index zero_row).  Lanes past
// the end of the segment carry zero_row as their index, so every trip gathers a full set of rows and the loop has no
// tail and no predication: short segments (17 entries on average on the drug graph) are otherwise dominated by the
// branchy remainder code.  `idx` = this lane's other[beg+lane], or zero_row past the end.
template <int LPR>
__device__ __forceinline__ float4 warp_gather_sum_padded(const float4* __restrict__ src, const int* __restrict__ other,
                                                         int beg, int end, int idx, int zero_row, int g, int l) {
    constexpr int G = 32 / LPR;
    constexpr int U = G >= 16 ? 32 / G : 4;  // rows per lane per trip; U * G entries per trip (a divisor of 32)
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = f4_zero();
    const float4* lane_src = src + l;
    for (int base = beg; base < end; base += 32) {
        if (base != beg) idx = (base + (threadIdx.x & 31) < end) ? ld_stream_i32(other + base + (threadIdx.x & 31)) : zero_row;
        const int cnt = min(32, end - base);
#pragma unroll 1
        for (int k0 = 0; k0 < cnt; k0 += U * G) {
            int j[U];
#pragma unroll
            for (int u = 0; u < U; ++u) j[u] = __shfl_sync(FULL, idx, k0 + u * G + g);
#pragma unroll
            for (int u = 0; u < U; ++u) acc[u] = f4_add(acc[u], lane_src[j[u] * LPR]);
        }
    }
    float4 a = acc[0];
#pragma unroll
    for (int u = 1; u < U; ++u) a = f4_add(a, acc[u]);
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        a.x += __shfl_xor_sync(FULL, a.x, o);
        a.y += __shfl_xor_sync(FULL, a.y, o);
        a.z += __shfl_xor_sync(FULL, a.z, o);
        a.w += __shfl_xor_sync(FULL, a.w, o);
    }
    return a;
}

template <int LPR, bool STAGED>
__global__ void __launch_bounds__(STAGED ? 1024 : 256)
k_seg_aggregate(const int* __restrict__ seg_ptr, const int* __restrict__ other, const int* __restrict__ counts,
                const float4* __restrict__ feat, const float* __restrict__ row_scale,
                const float4* __restrict__ relu_ref, int n_rows, float4* __restrict__ out) {
    extern __shared__ float4 s_feat[];
    const int lane = lane_id();
    const int g = lane / LPR, l = lane % LPR;

    if (STAGED) {
        for (int i = threadIdx.x; i < n_rows * LPR; i += blockDim.x) {
            float4 v = feat[i];
            if (relu_ref) {  // ReLU backward folded into the staging copy
                float4 r = relu_ref[i];
                v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
                v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
            }
            if (row_scale) {
                float sc = row_scale[i / LPR];
                v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            }
            s_feat[i] = v;
        }
        if (threadIdx.x < LPR) s_feat[n_rows * LPR + threadIdx.x] = f4_zero();  // the zero row
        __syncthreads();
    }
    const int pad = STAGED ? n_rows : 0;  // index carried by lanes past the end of a segment
    const float4* src = STAGED ? s_feat : feat;

    const int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= S) return;

    int beg = seg_ptr[s], end = seg_ptr[s + 1];
    int idx = (beg + lane < end) ? ld_stream_i32(other + beg + lane) : pad;
    while (true) {
        // prefetch the next segment of this warp
        const int sn = s + n_warps;
        int begn = 0, endn = 0, idxn = 0;
        if (sn < S) {
            begn = seg_ptr[sn];
            endn = seg_ptr[sn + 1];
            idxn = (begn + lane < endn) ? ld_stream_i32(other + begn + lane) : pad;
        }

        float4 a;
        if (STAGED) a = warp_gather_sum_padded<LPR>(src, other, beg, end, idx, pad, g, l);
        else a = warp_gather_sum<LPR, false>(src, other, row_scale, beg, end, idx);
        if (g == 0) out[int64_t(s) * LPR + l] = a;

        if (sn >= S) break;
        s = sn; beg = begn; end = endn; idx = idxn;
    }
}

// ------------------------------------------------------------------------------------------------
// Flat variant (feature matrix staged in shared memory, F in {16, 32, 64, 128}).
// Segments are short (17 entries on average on the drug graph), so a warp-per-segment loop spends most of its
// instructions on per-segment bookkeeping.  Here a warp owns a contiguous, entry-balanced range of segments and walks
// the ENTRY stream in 32-aligned blocks: all 32 lanes add the row of entry k (lane owns F/32 floats of the row), and
// when entry k is the last of its segment the accumulator is stored as that segment's row.  Segment ends inside a
// block come from 32 prefetched seg_ptr values and one warp OR-reduction; the next block's indices are prefetched.
// Lanes outside the warp's range carry the index of an all-zero row.
// acc[0..FPL) += the FPL floats at shared-memory byte address a
template <int FPL>
__device__ __forceinline__ void flat_add(float (&acc)[FPL], uint32_t a) {
    if (FPL == 1) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        acc[0] += v;
    } else if (FPL == 2) {
        float v0, v1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(a));
        const float2 t = __fadd2_rn(make_float2(acc[0], acc[1]), make_float2(v0, v1));
        acc[0] = t.x; acc[1] = t.y;
    } else {
        float v0, v1, v2, v3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(a));
        const float2 t0 = __fadd2_rn(make_float2(acc[0], acc[1]), make_float2(v0, v1));
        const float2 t1 = __fadd2_rn(make_float2(acc[2], acc[3]), make_float2(v2, v3));
        acc[0] = t0.x; acc[1] = t0.y; acc[2] = t1.x; acc[3] = t1.y;
    }
}

template <int F>
__global__ void __launch_bounds__(1024)
k_seg_aggregate_flat(const int* __restrict__ seg_ptr, const int* __restrict__ other, const int* __restrict__ counts,
                     const float4* __restrict__ feat, const float* __restrict__ row_scale,
                     const float4* __restrict__ relu_ref, int n_rows, float* __restrict__ out) {
    extern __shared__ float4 s_feat[];
    constexpr int LPR = F / 4;                     // float4 per row (staging)
    constexpr int FPL = F >= 32 ? F / 32 : 1;      // floats per lane
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_rows * LPR; i += blockDim.x) {
        float4 v = feat[i];
        if (relu_ref) {
            float4 r = relu_ref[i];
            v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
            v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
        }
        if (row_scale) {
            float sc = row_scale[i / LPR];
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        }
        s_feat[i] = v;
    }
    if (threadIdx.x < LPR) s_feat[n_rows * LPR + threadIdx.x] = f4_zero();
    __syncthreads();

    const int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    if (S <= 0) return;
    const int E = seg_ptr[S];
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int C = (E + n_warps - 1) / n_warps;
    int s, s_end;
    {   // first segment starting at or after w*C / (w+1)*C
        const int64_t t0 = int64_t(w) * C, t1 = t0 + C;
        const int a0 = t0 < E ? int(t0) : E, a1 = t1 < E ? int(t1) : E;
        int lo = 0, hi = S;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg_ptr[mid] < a0) lo = mid + 1; else hi = mid; }
        s = lo;
        hi = S;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg_ptr[mid] < a1) lo = mid + 1; else hi = mid; }
        s_end = lo;
    }
    if (s >= s_end) return;
    const int eb = seg_ptr[s], ee = seg_ptr[s_end];
    const int zero_row = n_rows;
    const uint32_t lane_addr = uint32_t(__cvta_generic_to_shared(s_feat)) + uint32_t((F >= 32 ? lane : (lane & (F - 1))) * FPL * 4);
    // the block's 32 row offsets are broadcast through a per-warp shared-memory slot (one STS, then uniform LDS.128:
    // four offsets per load) -- a SHFL per entry costs two LSU wavefronts and made the kernel LSU-pipe bound
    const uint32_t slot = uint32_t(__cvta_generic_to_shared(s_feat + (n_rows + 1) * LPR)) + uint32_t(threadIdx.x >> 5) * 128u;

    int b = eb & ~31;
    int idx, e_t;
    {
        const int p = b + lane;
        idx = (p >= eb && p < ee) ? ld_stream_i32(other + p) : zero_row;
        e_t = seg_ptr[min(s + 1 + lane, S)];
    }
    float acc[FPL];
#pragma unroll
    for (int q = 0; q < FPL; ++q) acc[q] = 0.f;

    for (; b < ee; b += 32) {
        const int q = e_t - 1 - b;
        unsigned flags = __reduce_or_sync(FULL, (q >= 0 && q < 32) ? 1u << q : 0u);
        const int lo = max(eb - b, 0), hi = min(ee - b, 32);
        flags &= (hi >= 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u);
        int idx_n = zero_row, e_n = 0;
        if (b + 32 < ee) {
            const int pn = b + 32 + lane;
            idx_n = pn < ee ? ld_stream_i32(other + pn) : zero_row;
            e_n = seg_ptr[min(s + __popc(flags) + 1 + lane, S)];
        }
        __syncwarp();
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot + uint32_t(lane) * 4u), "r"(uint32_t(idx) * uint32_t(F * 4)) : "memory");
        __syncwarp();
        // four entries per trip; a trip without a segment end (3 of 4 on the drug graph) takes the branch-free path:
        // predicating the store/reset on every entry costs more issue slots than the gather itself
#pragma unroll
        for (int k4 = 0; k4 < 32; k4 += 4) {
            uint32_t offs[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(offs[0]), "=r"(offs[1]), "=r"(offs[2]), "=r"(offs[3]) : "r"(slot + uint32_t(k4) * 4u));
            const unsigned m4 = (flags >> k4) & 15u;
            if (m4 == 0u) {
#pragma unroll
                for (int q = 0; q < 4; ++q) flat_add<FPL>(acc, lane_addr + offs[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    flat_add<FPL>(acc, lane_addr + offs[q]);
                    if (m4 & (1u << q)) {
                        float* dst = out + int64_t(s) * F + (F >= 32 ? lane : (lane & (F - 1))) * FPL;
                        if (FPL == 1) { if (F >= 32 || lane < F) dst[0] = acc[0]; }
                        else if (FPL == 2) *reinterpret_cast<float2*>(dst) = make_float2(acc[0], acc[1]);
                        else *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
                        for (int q2 = 0; q2 < FPL; ++q2) acc[q2] = 0.f;
                        ++s;
                    }
                }
            }
        }
        idx = idx_n;
        e_t = e_n;
    }
}

template <int F>
static int seg_aggregate_launch_flat(const CsrView& v, const float* feat, const float* row_scale,
                                     const float* relu_ref, int n_rows, float* out, cudaStream_t s) {
    const size_t bytes = size_t(n_rows + 1) * F * sizeof(float) + 32 * 128;  // rows + zero row + per-warp offset slots
    auto kern = k_seg_aggregate_flat<F>;
    if (int rc = ensure_dyn_smem((const void*)kern, bytes)) return rc;
    // two resident CTAs per SM when two copies of the feature matrix fit (F <= 32 on the drug graph)
    const int per_sm = 2 * (bytes + 1024) <= size_t(max_smem_optin()) ? 2 : 1;
    kern<<<sm_count() * per_sm, 1024, bytes, s>>>(v.seg_ptr, v.other, v.counts, (const float4*)feat, row_scale,
                                                  (const float4*)relu_ref, n_rows, out);
    TIPB_CHECK_LAUNCH("seg_aggregate_flat");
    return TIPB_OK;
}

// relu_ref is only honoured by the staged variant (callers fall back to pre-masking otherwise).
template <int LPR>
static int seg_aggregate_launch_lpr(const CsrView& v, const float* feat, const float* row_scale,
                                    const float* relu_ref, int n_rows, float* out, cudaStream_t s) {
    const size_t bytes = size_t(n_rows + 1) * LPR * sizeof(float4);  // + the zero row
    const bool staged = bytes + 1024 <= size_t(max_smem_optin());
    if (staged) {
        auto kern = k_seg_aggregate<LPR, true>;
        if (int rc = ensure_dyn_smem((const void*)kern, bytes)) return rc;
        kern<<<sm_count(), 1024, bytes, s>>>(v.seg_ptr, v.other, v.counts, (const float4*)feat, row_scale,
                                             (const float4*)relu_ref, n_rows, (float4*)out);
    } else {
        if (relu_ref) {
            set_last_error("seg_aggregate: relu_ref needs the staged path");
            return TIPB_ERR_UNSUPPORTED;
        }
        auto kern = k_seg_aggregate<LPR, false>;
        kern<<<sm_count() * 8, 256, 0, s>>>(v.seg_ptr, v.other, v.counts, (const float4*)feat, row_scale, nullptr,
                                            n_rows, (float4*)out);
    }
    TIPB_CHECK_LAUNCH("seg_aggregate");
    return TIPB_OK;
}

static inline bool seg_aggregate_supported(int f) { return f == 4 || f == 8 || f == 16 || f == 32 || f == 64 || f == 128; }

static int seg_aggregate_launch(const CsrView& v, const float* feat, const float* row_scale, const float* relu_ref,
                                int n_rows, int f, float* out, cudaStream_t s) {
    if (f >= 16 && size_t(n_rows + 1) * f * sizeof(float) + 32 * 128 + 1024 <= size_t(max_smem_optin())) {
        switch (f) {
            case 16: return seg_aggregate_launch_flat<16>(v, feat, row_scale, relu_ref, n_rows, out, s);
            case 32: return seg_aggregate_launch_flat<32>(v, feat, row_scale, relu_ref, n_rows, out, s);
            case 64: return seg_aggregate_launch_flat<64>(v, feat, row_scale, relu_ref, n_rows, out, s);
            case 128: return seg_aggregate_launch_flat<128>(v, feat, row_scale, relu_ref, n_rows, out, s);
        }
    }
    switch (f) {
        case 4: return seg_aggregate_launch_lpr<1>(v, feat, row_scale, relu_ref, n_rows, out, s);
        case 8: return seg_aggregate_launch_lpr<2>(v, feat, row_scale, relu_ref, n_rows, out, s);
        case 16: return seg_aggregate_launch_lpr<4>(v, feat, row_scale, relu_ref, n_rows, out, s);
        case 32: return seg_aggregate_launch_lpr<8>(v, feat, row_scale, relu_ref, n_rows, out, s);
        case 64: return seg_aggregate_launch_lpr<16>(v, feat, row_scale, relu_ref, n_rows, out, s);
        case 128: return seg_aggregate_launch_lpr<32>(v, feat, row_scale, relu_ref, n_rows, out, s);
    }
    set_last_error("seg_aggregate: feature width %d not in {4,8,16,32,64,128}", f);
    return TIPB_ERR_UNSUPPORTED;
}

}  // namespace tipb
