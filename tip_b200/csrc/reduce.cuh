// Relation-major reduction of per-segment rows (shared by the R-GCN d_att and the decoder d_weight).
#pragma once
#include "common.cuh"

namespace tipb {

// per-relation reduction of per-segment rows:  dst[r, :] = sum_{s in rel_seg[r]} src[s, :]   (row width w)
static __global__ void __launch_bounds__(128)
k_rel_reduce(const int* __restrict__ rel_seg_ptr, const int* __restrict__ rel_seg, const float* __restrict__ src,
             int w, float scale, float* __restrict__ dst) {
    __shared__ float part[4][128];
    const int r = blockIdx.x;
    const int beg = rel_seg_ptr[r], end = rel_seg_ptr[r + 1];
    const int wid = warp_id(), lane = lane_id();
    for (int c0 = 0; c0 < w; c0 += 32) {
        const int col = c0 + lane;
        float a0 = 0.f, a1 = 0.f;
        int p = beg + wid;
        for (; p + 4 < end; p += 8) {
            int s0 = rel_seg[p], s1 = rel_seg[p + 4];
            if (col < w) { a0 += src[int64_t(s0) * w + col]; a1 += src[int64_t(s1) * w + col]; }
        }
        for (; p < end; p += 4)
            if (col < w) a0 += src[int64_t(rel_seg[p]) * w + col];
        part[wid][lane] = a0 + a1;
        __syncthreads();
        if (wid == 0 && col < w)
            dst[int64_t(r) * w + col] = scale * (((part[0][lane] + part[1][lane]) + part[2][lane]) + part[3][lane]);
        __syncthreads();
    }
}

}  // namespace tipb
