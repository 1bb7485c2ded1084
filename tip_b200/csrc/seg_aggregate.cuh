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
        __syncthreads();
    }
    const float4* src = STAGED ? s_feat : feat;

    const int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= S) return;

    int beg = seg_ptr[s], end = seg_ptr[s + 1];
    int idx = (beg + lane < end) ? ld_stream_i32(other + beg + lane) : 0;
    while (true) {
        // prefetch the next segment of this warp
        const int sn = s + n_warps;
        int begn = 0, endn = 0, idxn = 0;
        if (sn < S) {
            begn = seg_ptr[sn];
            endn = seg_ptr[sn + 1];
            idxn = (begn + lane < endn) ? ld_stream_i32(other + begn + lane) : 0;
        }

        float4 a = warp_gather_sum<LPR, STAGED>(src, other, row_scale, beg, end, idx);
        if (g == 0) out[int64_t(s) * LPR + l] = a;

        if (sn >= S) break;
        s = sn; beg = begn; end = endn; idx = idxn;
    }
}

// relu_ref is only honoured by the staged variant (callers fall back to pre-masking otherwise).
template <int LPR>
static int seg_aggregate_launch_lpr(const CsrView& v, const float* feat, const float* row_scale,
                                    const float* relu_ref, int n_rows, float* out, cudaStream_t s) {
    const size_t bytes = size_t(n_rows) * LPR * sizeof(float4);
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
