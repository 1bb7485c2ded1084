// Relation-major reduction of per-segment rows (shared by the R-GCN d_att and the decoder d_weight).
#pragma once
#include "common.cuh"

namespace tipb {

constexpr int REL_REDUCE_THREADS = 256;

// dst[r, :] = scale * sum_{s in rel_seg[rel_seg_ptr[r] .. rel_seg_ptr[r+1])} src[s, :]      (row width w)
// One CTA per relation; warp k takes rows k, k+8, ... with four rows in flight; the eight warp partials are
// added in a fixed order, so the result does not depend on scheduling.
static __global__ void __launch_bounds__(REL_REDUCE_THREADS)
k_rel_reduce(const int* __restrict__ rel_seg_ptr, const int* __restrict__ rel_seg, const int* __restrict__ counts,
             const float* __restrict__ src, int w, float scale, float* __restrict__ dst) {
    constexpr int NW = REL_REDUCE_THREADS / 32;
    __shared__ float part[NW][32];
    const int r = blockIdx.x;
    const int beg = rel_seg_ptr[r], end = rel_seg_ptr[r + 1];
    const bool direct = counts[TIPB_CSR_COUNT_REL_MAJOR] != 0;  // relation-major plan: rows beg..end are the segments
    const int wid = warp_id(), lane = lane_id();
    for (int c0 = 0; c0 < w; c0 += 32) {
        const int col = c0 + lane;
        const bool ok = col < w;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int p = beg + wid;
        for (; p + 3 * NW < end; p += 4 * NW) {
            const int s0 = direct ? p : rel_seg[p], s1 = direct ? p + NW : rel_seg[p + NW];
            const int s2 = direct ? p + 2 * NW : rel_seg[p + 2 * NW], s3 = direct ? p + 3 * NW : rel_seg[p + 3 * NW];
            if (ok) {
                a0 += src[int64_t(s0) * w + col];
                a1 += src[int64_t(s1) * w + col];
                a2 += src[int64_t(s2) * w + col];
                a3 += src[int64_t(s3) * w + col];
            }
        }
        for (; p < end; p += NW)
            if (ok) a0 += src[int64_t(direct ? p : rel_seg[p]) * w + col];
        part[wid][lane] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (wid == 0 && ok) {
            float t = part[0][lane];
#pragma unroll
            for (int k = 1; k < NW; ++k) t += part[k][lane];
            dst[int64_t(r) * w + col] = scale * t;
        }
        __syncthreads();
    }
}

}  // namespace tipb
