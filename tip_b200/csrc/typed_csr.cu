// GPU-side construction of the "typed CSR" every graph kernel in this library consumes
// (north_star item 6: CSR-by-relation construction on the GPU).
//
// Input is the reference's edge layout: int64 edge_index [2,E] plus either
// an int64 edge_type [E] (MyRGCNConv, src/layers.py:76-79: any order) or the int64 range_list [R,2]
// of relation-sorted edges (MyRGCNConv2, src/layers.py:157-160; src/utils.py:26-32).
//
// Output ("plan", one caller-owned device buffer; field offsets from tipb_typed_csr_layout):
//   entries are grouped by (node, relation) -- node = target endpoint (by_src=0) or source
//   endpoint (by_src=1) -- and keep their input order inside a group (stable sort), so every
//   downstream floating-point sum has ONE fixed order: no atomics, run-to-run deterministic.
//     eid[p]      input position of the p-th entry            other[p]   its opposite endpoint
//     seg_ptr[s]  first entry of non-empty segment s          seg_node/seg_rel[s]
//     node_ptr[n] first segment of node n                     deg/inv_deg[n]  = #entries, 1/max(#,1)
//     rel_seg_ptr[r], rel_seg[]  the same segments listed relation-major (for per-relation reductions)
//   `doubled=1` lists every edge in both directions (2E entries; entry e+E is edge e reversed): the
//   decoder's symmetric score makes d(loss)/dz a sum over both endpoints of every pair.
#include "common.cuh"

namespace tipb {

struct CsrLayout {
    int64_t entries, n_nodes, n_rel, seg_cap;
    int64_t off[TIPB_CSR_NFIELDS];
    int64_t total_bytes;
};

static CsrLayout make_layout(int64_t entries, int64_t n_nodes, int64_t n_rel) {
    CsrLayout L;
    L.entries = entries;
    L.n_nodes = n_nodes;
    L.n_rel = n_rel;
    int64_t cap = n_nodes * n_rel;
    L.seg_cap = entries < cap ? entries : cap;
    if (L.seg_cap < 1) L.seg_cap = 1;
    int64_t counts[TIPB_CSR_NFIELDS];
    counts[TIPB_CSR_COUNTS] = 16;
    counts[TIPB_CSR_EID] = entries;
    counts[TIPB_CSR_OTHER] = entries;
    counts[TIPB_CSR_SEG_PTR] = L.seg_cap + 1;
    counts[TIPB_CSR_SEG_NODE] = L.seg_cap;
    counts[TIPB_CSR_SEG_REL] = L.seg_cap;
    counts[TIPB_CSR_NODE_PTR] = n_nodes + 1;
    counts[TIPB_CSR_DEG] = n_nodes;
    counts[TIPB_CSR_INV_DEG] = n_nodes;
    counts[TIPB_CSR_REL_SEG_PTR] = n_rel + 1;
    counts[TIPB_CSR_REL_SEG] = L.seg_cap;
    int64_t off = 0;
    for (int f = 0; f < TIPB_CSR_NFIELDS; ++f) {
        L.off[f] = off;
        off += ((counts[f] > 0 ? counts[f] : 1) * 4 + 255) & ~int64_t(255);
    }
    L.total_bytes = off;
    return L;
}

CsrView csr_view(const void* plan, int64_t entries, int64_t n_nodes, int64_t n_rel) {
    CsrLayout L = make_layout(entries, n_nodes, n_rel);
    const char* b = static_cast<const char*>(plan);
    CsrView v;
    v.entries = entries;
    v.n_nodes = n_nodes;
    v.n_rel = n_rel;
    v.seg_cap = L.seg_cap;
    v.counts = (int*)(b + L.off[TIPB_CSR_COUNTS]);
    v.eid = (int*)(b + L.off[TIPB_CSR_EID]);
    v.other = (int*)(b + L.off[TIPB_CSR_OTHER]);
    v.seg_ptr = (int*)(b + L.off[TIPB_CSR_SEG_PTR]);
    v.seg_node = (int*)(b + L.off[TIPB_CSR_SEG_NODE]);
    v.seg_rel = (int*)(b + L.off[TIPB_CSR_SEG_REL]);
    v.node_ptr = (int*)(b + L.off[TIPB_CSR_NODE_PTR]);
    v.deg = (int*)(b + L.off[TIPB_CSR_DEG]);
    v.inv_deg = (float*)(b + L.off[TIPB_CSR_INV_DEG]);
    v.rel_seg_ptr = (int*)(b + L.off[TIPB_CSR_REL_SEG_PTR]);
    v.rel_seg = (int*)(b + L.off[TIPB_CSR_REL_SEG]);
    return v;
}

// segment key: node-major plans sort by (node, relation), relation-major plans by (relation, node)
__host__ __device__ __forceinline__ int64_t seg_key(int64_t node, int64_t rel, int n_nodes, int n_rel, int rel_major) {
    return rel_major ? rel * n_nodes + node : node * n_rel + rel;
}
__host__ __device__ __forceinline__ void seg_unkey(uint32_t key, int n_nodes, int n_rel, int rel_major, int& node,
                                                   int& rel) {
    if (rel_major) { rel = int(key / uint32_t(n_nodes)); node = int(key % uint32_t(n_nodes)); }
    else { node = int(key / uint32_t(n_rel)); rel = int(key % uint32_t(n_rel)); }
}

// ------------------------------------------------------------------------------------------------
// 1. keys: key = seg_key(node, rel) (sentinel N*R for dropped / out-of-range entries), val = entry id
__global__ void k_csr_keys(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ edge_type,
                           const int64_t* __restrict__ range_list, int64_t E, int64_t entries, int n_nodes,
                           int n_other, int n_rel, int by_src, int drop_loops, int rel_major,
                           uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ counts) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= entries) return;
    int64_t e = p < E ? p : p - E;
    bool rev = p >= E;
    int64_t a = edge_index[e], b = edge_index[E + e];  // a = row 0 (source), b = row 1 (target)
    int64_t node = (by_src != 0) != rev ? a : b;
    int64_t other = (by_src != 0) != rev ? b : a;
    int64_t rel = 0;
    if (edge_type) {
        rel = edge_type[e];
    } else if (range_list) {
        // relation whose [start,end) holds e: last r with start <= e (ranges are cumulative)
        int lo = 0, hi = n_rel - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (range_list[2 * mid] <= e) lo = mid; else hi = mid - 1;
        }
        rel = lo;
        // skip over empty trailing ranges that share the same start
        if (!(range_list[2 * rel] <= e && e < range_list[2 * rel + 1])) {
            rel = -1;
            for (int r = lo; r >= 0 && range_list[2 * r] == range_list[2 * lo]; --r)
                if (e < range_list[2 * r + 1]) { rel = r; break; }
        }
    }
    bool ok = node >= 0 && node < n_nodes && other >= 0 && other < n_other && rel >= 0 && rel < n_rel;
    if (!ok) atomicOr(&counts[TIPB_CSR_COUNT_STATUS], 1);
    bool keep = ok && !(drop_loops && node == other);
    keys[p] = keep ? uint32_t(seg_key(node, rel, n_nodes, n_rel, rel_major)) : uint32_t(int64_t(n_nodes) * n_rel);
    vals[p] = uint32_t(p);
}

// 2. after the sort: other endpoint per entry + segment-start flags
__global__ void k_csr_flags(const int64_t* __restrict__ edge_index, int64_t E, int64_t entries, int by_src,
                            uint32_t sentinel, const uint32_t* __restrict__ keys, const int* __restrict__ eid,
                            int* __restrict__ other, int* __restrict__ flags) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= entries) return;
    uint32_t k = keys[p];
    int id = eid[p];
    int64_t e = id < E ? id : id - E;
    bool rev = id >= E;
    int row = ((by_src != 0) != rev) ? 1 : 0;  // row holding the OTHER endpoint
    other[p] = k == sentinel ? 0 : int(edge_index[int64_t(row) * E + e]);
    flags[p] = (k != sentinel && (p == 0 || keys[p - 1] != k)) ? 1 : 0;
}

// 3. segment table from the scanned flags
__global__ void k_csr_segments(int64_t entries, uint32_t sentinel, int n_nodes, int n_rel, int rel_major,
                               const uint32_t* __restrict__ keys, const int* __restrict__ seg_index /*excl scan*/,
                               int* __restrict__ seg_ptr, int* __restrict__ seg_node, int* __restrict__ seg_rel,
                               int* __restrict__ counts) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p > entries) return;
    if (p == entries) {  // one past: close the last segment, publish S
        int S = seg_index[entries];
        counts[TIPB_CSR_COUNT_SEGMENTS] = S;
        return;
    }
    uint32_t k = keys[p];
    bool start = k != sentinel && (p == 0 || keys[p - 1] != k);
    if (start) {
        int s = seg_index[p];
        seg_ptr[s] = int(p);
        seg_unkey(k, n_nodes, n_rel, rel_major, seg_node[s], seg_rel[s]);
    }
    // first dropped entry (or the end) terminates the valid range
    bool last_valid = k != sentinel && (p + 1 == entries || keys[p + 1] == sentinel);
    if (last_valid) {
        int S = seg_index[p] + (start ? 1 : 0);
        // seg_index is exclusive: segments started strictly before p, plus p's own start
        seg_ptr[S] = int(p + 1);
        counts[TIPB_CSR_COUNT_VALID] = int(p + 1);
    }
    if (p == 0 && k == sentinel) {  // nothing valid at all
        seg_ptr[0] = 0;
        counts[TIPB_CSR_COUNT_VALID] = 0;
    }
}

// pad the unused tail of the per-segment tables with sentinels so host-sized launches stay simple
// (rkeys, rvals) = keys of the SECONDARY listing: relation for node-major plans, node for relation-major plans
__global__ void k_csr_pad(int64_t seg_cap, int n_nodes, int n_rel, int rel_major, int* __restrict__ counts,
                          int* __restrict__ seg_node, int* __restrict__ seg_rel, uint32_t* __restrict__ rkeys,
                          uint32_t* __restrict__ rvals) {
    int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= seg_cap) return;
    if (s == 0) counts[TIPB_CSR_COUNT_REL_MAJOR] = rel_major;
    int S = counts[TIPB_CSR_COUNT_SEGMENTS];
    if (s >= S) {
        seg_node[s] = n_nodes;
        seg_rel[s] = n_rel;
    }
    rkeys[s] = uint32_t(rel_major ? seg_node[s] : seg_rel[s]);
    rvals[s] = uint32_t(s);
}

// degrees of a relation-major plan: a node's segments are reached through the secondary listing (one warp per node)
__global__ void __launch_bounds__(256)
k_csr_degrees_listed(int n_nodes, const int* __restrict__ node_ptr, const int* __restrict__ listing,
                     const int* __restrict__ seg_ptr, int* __restrict__ deg, float* __restrict__ inv_deg) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_nodes) return;
    int d = 0;
    for (int i = node_ptr[n] + lane_id(); i < node_ptr[n + 1]; i += 32) {
        const int s = listing[i];
        d += seg_ptr[s + 1] - seg_ptr[s];
    }
    d = warp_sum_i(d);
    if (lane_id() == 0) {
        deg[n] = d;
        inv_deg[n] = 1.0f / float(d < 1 ? 1 : d);
    }
}

// group_ptr[g] = first index i with sorted_keys[i] >= g, for g in [0, n_groups]
template <typename K>
__global__ void k_group_ptr(const K* __restrict__ sorted_keys, int64_t n, int n_groups, int* __restrict__ ptr) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t prev = i == 0 ? -1 : int64_t(sorted_keys[i - 1]);
    int64_t cur = i == n ? int64_t(n_groups) : int64_t(sorted_keys[i]);
    if (prev > n_groups) prev = n_groups;
    if (cur > n_groups) cur = n_groups;
    for (int64_t g = prev + 1; g <= cur; ++g) ptr[g] = int(i);
}

__global__ void k_csr_degrees(int n_nodes, const int* __restrict__ node_ptr, const int* __restrict__ seg_ptr,
                              int* __restrict__ deg, float* __restrict__ inv_deg) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    int d = seg_ptr[node_ptr[n + 1]] - seg_ptr[node_ptr[n]];
    deg[n] = d;
    inv_deg[n] = 1.0f / float(d < 1 ? 1 : d);
}

__global__ void k_copy_u32_to_i32(const uint32_t* __restrict__ in, int* __restrict__ out, int64_t n) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = int(in[i]);
}

static int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 32 && (uint64_t(1) << b) <= max_value) ++b;
    return b;
}

// ================================================================================================
// Grouped build: when the edges arrive sorted by relation (range_list given -- the layout of
// src/utils.py:35-65, and of every per-step negative sample), no global sort is needed.  The
// (node, relation) segment sizes come from one histogram over the dense key space, their offsets
// from one scan, and each relation is then placed by ONE CTA that walks its edges in input order
// (per-warp counters + a per-node cursor in shared memory), which keeps the placement stable and
// therefore bit-identical to the sort-based path.
// ================================================================================================
__device__ __forceinline__ bool grp_entry(const int64_t* __restrict__ edge_index, int64_t E, int64_t e, bool rev,
                                          int by_src, int drop_loops, int n_nodes, int n_other, int& node,
                                          int& other) {
    const int64_t a = edge_index[e], b = edge_index[E + e];
    const bool pick_a = (by_src != 0) != rev;
    const int64_t nd = pick_a ? a : b, ot = pick_a ? b : a;
    node = int(nd);
    other = int(ot);
    return nd >= 0 && nd < n_nodes && ot >= 0 && ot < n_other && !(drop_loops && nd == ot);
}

// status bit1: range_list is not the cumulative cover of [0, E)
__global__ void k_grp_check_ranges(const int64_t* __restrict__ range_list, int n_rel, int64_t E, int* __restrict__ counts) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rel) return;
    const int64_t s = range_list[2 * r], e = range_list[2 * r + 1];
    const int64_t prev_end = r == 0 ? 0 : range_list[2 * r - 1];
    bool ok = s == prev_end && e >= s && (r + 1 < n_rel || e == E);
    if (!ok) atomicOr(&counts[TIPB_CSR_COUNT_STATUS], 2);
}

// cnt[node*R + r] = entries of segment (node, r); cnt_fwd = the forward-direction share (doubled plans place the
// reversed copies after the forward ones, so the second pass starts its per-node cursor there)
__global__ void __launch_bounds__(256)
k_grp_count(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ range_list, int64_t E, int doubled,
            int n_nodes, int n_other, int n_rel, int by_src, int drop_loops, int rel_major, int* __restrict__ cnt,
            int* __restrict__ cnt_fwd, int* __restrict__ counts) {
    extern __shared__ int s_hist[];  // [2][n_nodes]: both directions, forward only
    const int r = blockIdx.y;
    const int64_t start = range_list[2 * r], end = range_list[2 * r + 1];
    for (int n = threadIdx.x; n < 2 * n_nodes; n += blockDim.x) s_hist[n] = 0;
    __syncthreads();
    bool bad = false;
    // four independent pairs of 64-bit loads per thread and trip: the pass is a pure stream over the relation
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e0 = start + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e0 < end; e0 += 4 * stride) {
        int64_t ra[4], rb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t e = e0 + j * stride;
            const int64_t ec = e < end ? e : end - 1;
            ra[j] = edge_index[ec];
            rb[j] = edge_index[E + ec];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (e0 + j * stride >= end) break;
            const int64_t nd = by_src ? ra[j] : rb[j], ot = by_src ? rb[j] : ra[j];
            const bool in = nd >= 0 && nd < n_nodes && ot >= 0 && ot < n_other;
            const bool loop = drop_loops && nd == ot;
            if (in && !loop) {
                atomicAdd(&s_hist[int(nd)], 1);
                atomicAdd(&s_hist[n_nodes + int(nd)], 1);
                // the reversed copy of a doubled plan lists the pair under its other endpoint
                if (doubled) atomicAdd(&s_hist[int(ot)], 1);
            } else if (!loop) {
                bad = true;
            }
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < n_nodes; n += blockDim.x) {
        const int64_t key = seg_key(n, r, n_nodes, n_rel, rel_major);
        if (s_hist[n]) atomicAdd(&cnt[key], s_hist[n]);
        if (doubled && s_hist[n_nodes + n]) atomicAdd(&cnt_fwd[key], s_hist[n_nodes + n]);
    }
    if (bad) atomicOr(&counts[TIPB_CSR_COUNT_STATUS], 1);
}

constexpr int GRP_ROUNDS = 8;

// One CTA per (relation, direction).  A tile is WARPS x 16 rounds x 32 consecutive edges; every thread first
// pulls its 16 edges into registers (one memory round trip), then the tile is ranked twice over the registers:
// per-warp per-node counts -> prefix over warps + running per-node cursor -> stable positions.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 3 : 1)
k_grp_scatter(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ range_list, int64_t E,
              int n_nodes, int n_other, int n_rel, int by_src, int drop_loops, int rel_major,
              const int* __restrict__ seg_start, const int* __restrict__ cnt_fwd, int* __restrict__ eid,
              int* __restrict__ other_out) {
    extern __shared__ int sm_i[];
    int* wh = sm_i;                        // [WARPS][n_nodes]
    int* cursor = sm_i + WARPS * n_nodes;  // [n_nodes]
    const int r = blockIdx.x, dir = blockIdx.y;
    const int w = warp_id(), lane = lane_id();
    const int64_t start = range_list[2 * r], end = range_list[2 * r + 1];
    for (int n = threadIdx.x; n < n_nodes; n += WARPS * 32)
        cursor[n] = dir ? cnt_fwd[seg_key(n, r, n_nodes, n_rel, rel_major)] : 0;
    constexpr int TILE = WARPS * 32 * GRP_ROUNDS;
    for (int64_t tile = start; tile < end; tile += TILE) {
        for (int i = threadIdx.x; i < WARPS * n_nodes; i += WARPS * 32) wh[i] = 0;
        const int64_t wbase = tile + int64_t(w) * 32 * GRP_ROUNDS;
        int node[GRP_ROUNDS], oth[GRP_ROUNDS], base[GRP_ROUNDS];
        unsigned vmask = 0;
        // two batches of eight rounds: all sixteen 64-bit loads of a batch are issued before any is consumed
        // (clamped addresses keep them unconditional), so a tile costs two memory round trips, not sixteen
        const bool pick_a = (by_src != 0) != (dir != 0);
#pragma unroll
        for (int half = 0; half < GRP_ROUNDS; half += 8) {
            int64_t ra[8], rb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t e = wbase + (half + j) * 32 + lane;
                const int64_t ec = e < end ? e : end - 1;
                ra[j] = edge_index[ec];
                rb[j] = edge_index[E + ec];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t e = wbase + (half + j) * 32 + lane;
                const int64_t nd = pick_a ? ra[j] : rb[j], ot = pick_a ? rb[j] : ra[j];
                node[half + j] = int(nd);
                oth[half + j] = int(ot);
                if (e < end && nd >= 0 && nd < n_nodes && ot >= 0 && ot < n_other && !(drop_loops && nd == ot))
                    vmask |= 1u << (half + j);
            }
        }
#pragma unroll
        for (int rd = 0; rd < GRP_ROUNDS; ++rd)
            base[rd] = (vmask >> rd) & 1u ? seg_start[seg_key(node[rd], r, n_nodes, n_rel, rel_major)] : 0;
        __syncthreads();
#pragma unroll
        for (int rd = 0; rd < GRP_ROUNDS; ++rd) {
            const bool valid = (vmask >> rd) & 1u;
            const unsigned act = __ballot_sync(FULL, valid);
            if (valid) {
                const unsigned peers = __match_any_sync(act, node[rd]);
                if ((__ffs(peers) - 1) == lane) wh[w * n_nodes + node[rd]] += __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        for (int n = threadIdx.x; n < n_nodes; n += WARPS * 32) {
            int run = cursor[n];
#pragma unroll
            for (int ww = 0; ww < WARPS; ++ww) {
                const int c = wh[ww * n_nodes + n];
                wh[ww * n_nodes + n] = run;
                run += c;
            }
            cursor[n] = run;
        }
        __syncthreads();
#pragma unroll
        for (int rd = 0; rd < GRP_ROUNDS; ++rd) {
            const bool valid = (vmask >> rd) & 1u;
            const unsigned act = __ballot_sync(FULL, valid);
            unsigned peers = 0;
            int pos = 0;
            if (valid) {
                peers = __match_any_sync(act, node[rd]);
                pos = base[rd] + wh[w * n_nodes + node[rd]] + __popc(peers & ((1u << lane) - 1u));
            }
            __syncwarp();
            if (valid && (__ffs(peers) - 1) == lane) wh[w * n_nodes + node[rd]] += __popc(peers);
            __syncwarp();
            if (valid) {
                const int64_t e = wbase + rd * 32 + lane;
                eid[pos] = int(dir ? E + e : e);
                other_out[pos] = oth[rd];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Fast placement (node and other ids < 65535).  One CTA per (relation, direction), handed out largest relation
// first (`order`), walks the relation in tiles of PLACE_TILE edges staged in shared memory as 16-bit ids:
//   count    lanes holding the same node find each other with one ballot per key bit (__match_any_sync costs
//            ~12 cycles per distinct value on sm_100a); warp w owns a contiguous run of the tile, so
//            (warp, round, lane) order is input order; each entry keeps its rank inside its warp's run
//   scan     per node: prefix over warps, then an exclusive scan over nodes -> the tile sorted by node, locally
//   place    sorted_i[local position] = tile index (a shared-memory scatter)
//   write    thread j writes local position j: consecutive threads write consecutive entries of a segment, so
//            the global stores coalesce (a direct scatter touches one sector per entry).
constexpr int PLACE_TILE = 2048;
constexpr uint16_t PLACE_INVALID = 0xffffu;

// order[rank] = relation, by decreasing size (ties by index): longest-processing-time-first over the CTAs
__global__ void __launch_bounds__(256) k_rel_order(const int64_t* __restrict__ range_list, int n_rel,
                                                   int* __restrict__ order) {
    __shared__ int s_size[1024];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int mine = r < n_rel ? int(range_list[2 * r + 1] - range_list[2 * r]) : 0;
    int rank = 0;
    for (int q0 = 0; q0 < n_rel; q0 += 1024) {
        __syncthreads();
        for (int q = threadIdx.x; q < 1024; q += blockDim.x)
            s_size[q] = q0 + q < n_rel ? int(range_list[2 * (q0 + q) + 1] - range_list[2 * (q0 + q)]) : -1;
        __syncthreads();
        const int lim = min(1024, n_rel - q0);
        for (int q = 0; q < lim; ++q) {
            const int sz = s_size[q];
            rank += (sz > mine || (sz == mine && q0 + q < r)) ? 1 : 0;
        }
    }
    if (r < n_rel) order[rank] = r;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_grp_place(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ range_list,
            const int* __restrict__ order, int64_t E, int n_nodes, int n_other, int n_rel, int by_src, int drop_loops,
            int rel_major, int doubled, int key_bits, const int* __restrict__ seg_start,
            const int* __restrict__ cnt_fwd, int* __restrict__ eid, int* __restrict__ other_out) {
    extern __shared__ int sm_i[];
    constexpr int T = WARPS * 32;
    int* cursor = sm_i;                                   // [n_nodes] absolute position of a node's next entry
    int* delta = cursor + n_nodes;                        // [n_nodes] absolute - local position, this tile
    int* lstart = delta + n_nodes;                        // [n_nodes] local position of a node's first entry
    uint16_t* wh = reinterpret_cast<uint16_t*>(lstart + n_nodes);  // [WARPS][n_nodes]
    uint16_t* cn = wh + ((WARPS * n_nodes + 1) & ~1);     // [TILE] node id (PLACE_INVALID: dropped entry)
    uint16_t* co = cn + PLACE_TILE;                       // [TILE] other id
    uint16_t* rk = co + PLACE_TILE;                       // [TILE] rank inside the warp's run
    uint16_t* sorted_i = rk + PLACE_TILE;                 // [TILE] tile index by local position
    __shared__ int s_scan[WARPS + 1];

    const int dirs = doubled ? 2 : 1;
    const int r = order[blockIdx.x / dirs], dir = blockIdx.x % dirs;
    const int w = warp_id(), lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const int64_t start = range_list[2 * r], end = range_list[2 * r + 1];
    for (int n = threadIdx.x; n < n_nodes; n += T) {
        const int64_t key = seg_key(n, r, n_nodes, n_rel, rel_major);
        cursor[n] = seg_start[key] + (dir ? cnt_fwd[key] : 0);
    }
    const bool node_is_a = (by_src != 0) != (dir != 0);  // column 0 (source) is this pass's node column
    const int per = (n_nodes + T - 1) / T;               // nodes per thread in the scan phase
    uint16_t* mywh = wh + w * n_nodes;
    for (int64_t tile = start; tile < end; tile += PLACE_TILE) {
        const int nt = int(end - tile < PLACE_TILE ? end - tile : PLACE_TILE);
        __syncthreads();  // previous tile written out
        for (int i0 = threadIdx.x; i0 < nt; i0 += 4 * T) {
            int64_t ra[4], rb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * T;
                const int64_t e = tile + (i < nt ? i : nt - 1);
                ra[j] = edge_index[e];
                rb[j] = edge_index[E + e];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * T;
                if (i < nt) {
                    const int64_t nd = node_is_a ? ra[j] : rb[j], ot = node_is_a ? rb[j] : ra[j];
                    const bool ok = nd >= 0 && nd < n_nodes && ot >= 0 && ot < n_other && !(drop_loops && nd == ot);
                    cn[i] = ok ? uint16_t(nd) : PLACE_INVALID;
                    co[i] = uint16_t(ot);
                }
            }
        }
        for (int i = threadIdx.x; i < (WARPS * n_nodes + 1) / 2; i += T) reinterpret_cast<uint32_t*>(wh)[i] = 0u;
        __syncthreads();
        // ---- count
        const int rounds = (((nt + WARPS - 1) / WARPS) + 31) >> 5;  // per warp
        const int wbase = w * rounds * 32;
        for (int rd = 0; rd < rounds; ++rd) {
            const int i = wbase + rd * 32 + lane;
            const uint16_t v = i < nt ? cn[i] : PLACE_INVALID;
            const bool valid = v != PLACE_INVALID;
            const int node = v;
            unsigned peers = __ballot_sync(FULL, valid);
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                if (b < key_bits) {
                    const bool bit = (node >> b) & 1;
                    const unsigned m = __ballot_sync(FULL, bit);
                    peers &= bit ? m : ~m;
                }
            }
            int c0 = 0;
            if (valid) c0 = mywh[node];
            __syncwarp();
            if (valid) {
                const int rank = __popc(peers & lt);
                if (rank == 0) mywh[node] = uint16_t(c0 + __popc(peers));
                rk[i] = uint16_t(c0 + rank);
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- scan: thread t owns nodes [t*per, (t+1)*per)
        int mine = 0;
        for (int j = 0; j < per; ++j) {
            const int n = threadIdx.x * per + j;
            if (n < n_nodes) {
                int run = 0;
#pragma unroll
                for (int ww = 0; ww < WARPS; ++ww) {
                    const int c = wh[ww * n_nodes + n];
                    wh[ww * n_nodes + n] = uint16_t(run);
                    run += c;
                }
                lstart[n] = run;  // the node's count for now
                mine += run;
            }
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_scan[w] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int ww = 0; ww < WARPS; ++ww) {
                const int c = s_scan[ww];
                s_scan[ww] = run;
                run += c;
            }
            s_scan[WARPS] = run;
        }
        __syncthreads();
        int ls = s_scan[w] + incl - mine;
        for (int j = 0; j < per; ++j) {
            const int n = threadIdx.x * per + j;
            if (n < n_nodes) {
                const int c = lstart[n];
                lstart[n] = ls;
                delta[n] = cursor[n] - ls;
                cursor[n] += c;
                ls += c;
            }
        }
        const int n_valid = s_scan[WARPS];
        __syncthreads();
        // ---- place (shared memory)
        for (int rd = 0; rd < rounds; ++rd) {
            const int i = wbase + rd * 32 + lane;
            if (i < nt) {
                const int node = cn[i];
                if (node != PLACE_INVALID) sorted_i[lstart[node] + mywh[node] + rk[i]] = uint16_t(i);
            }
        }
        __syncthreads();
        // ---- write, coalesced
        for (int j = threadIdx.x; j < n_valid; j += T) {
            const int i = sorted_i[j];
            const int pos = j + delta[cn[i]];
            const int64_t e = tile + i;
            eid[pos] = int(dir ? E + e : e);
            other_out[pos] = int(co[i]);
        }
    }
}

static size_t place_smem(int warps, int64_t n_nodes) {
    return size_t(3) * n_nodes * 4 + ((size_t(warps) * n_nodes + 1) & ~size_t(1)) * 2 + size_t(4) * PLACE_TILE * 2;
}
static int place_warps(int64_t n_nodes, int64_t n_other) {
    if (n_nodes > 65534 || n_other > 65535) return 0;
    if (place_smem(8, n_nodes) <= 200 * 1024) return 8;
    if (place_smem(4, n_nodes) <= 200 * 1024) return 4;
    return 0;
}

// dense key space -> compact segment table.  Keys are q = a * n_b + b with (a, b) = (relation, node) for
// relation-major plans and (node, relation) for node-major ones.  The SECONDARY listing (the same segments grouped
// by b) needs no sort either: scanning the non-empty flags in transposed key order (b * n_a + a) gives every
// segment its position in that listing.
__global__ void k_grp_flags(const int* __restrict__ cnt, int64_t n_keys, int n_a, int n_b, int* __restrict__ flags,
                            int* __restrict__ flags_t) {
    int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= n_keys) return;
    const int f = cnt[q] > 0 ? 1 : 0;
    flags[q] = f;
    const int64_t a = q / n_b, b = q % n_b;
    flags_t[b * n_a + a] = f;
}

__global__ void k_grp_segments(const int* __restrict__ cnt, const int* __restrict__ seg_start,
                               const int* __restrict__ seg_index, const int* __restrict__ pos_t, int64_t n_keys,
                               int n_nodes, int n_rel, int rel_major, int* __restrict__ seg_ptr,
                               int* __restrict__ seg_node, int* __restrict__ seg_rel, int* __restrict__ listing,
                               int* __restrict__ counts) {
    int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q > n_keys) return;
    if (q == n_keys) {
        const int S = seg_index[n_keys];
        counts[TIPB_CSR_COUNT_SEGMENTS] = S;
        counts[TIPB_CSR_COUNT_VALID] = seg_start[n_keys];
        counts[TIPB_CSR_COUNT_REL_MAJOR] = rel_major;
        seg_ptr[S] = seg_start[n_keys];
        return;
    }
    if (cnt[q] > 0) {
        const int s = seg_index[q];
        seg_ptr[s] = seg_start[q];
        int node, rel;
        seg_unkey(uint32_t(q), n_nodes, n_rel, rel_major, node, rel);
        seg_node[s] = node;
        seg_rel[s] = rel;
        const int64_t tq = rel_major ? int64_t(node) * n_rel + rel : int64_t(rel) * n_nodes + node;
        listing[pos_t[tq]] = s;
    }
}

// group pointers of both orders straight from the two scans, and the sentinel padding of the unused table tail
__global__ void k_grp_tail(const int* __restrict__ seg_index, const int* __restrict__ pos_t, int64_t seg_cap, int n_a,
                           int n_b, int n_nodes, int n_rel, int* __restrict__ ptr_a, int* __restrict__ ptr_b,
                           int* __restrict__ seg_node, int* __restrict__ seg_rel, int* __restrict__ listing) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i <= n_a) ptr_a[i] = seg_index[i * n_b];
    if (i <= n_b) ptr_b[i] = pos_t[i * n_a];
    const int S = seg_index[int64_t(n_a) * n_b];
    if (i >= S && i < seg_cap) {
        seg_node[i] = n_nodes;
        seg_rel[i] = n_rel;
        listing[i] = int(i);
    }
}

static int grp_warps(int64_t n_nodes) {
    // shared memory: (WARPS + 1) * n_nodes ints
    if (9 * n_nodes * 4 <= 96 * 1024) return 8;
    if (5 * n_nodes * 4 <= 160 * 1024) return 4;
    if (3 * n_nodes * 4 <= 200 * 1024) return 2;
    return 0;
}

static bool use_grouped(const int64_t* edge_type, const int64_t* range_list, int64_t entries, int64_t n_nodes,
                        int64_t n_other, int64_t n_rel, int doubled) {
    // (a doubled plan lists every pair under both endpoints: one validity rule for both copies needs one id space)
    // (large node counts -- 10^4 drugs -- have no fast in-CTA placement: per-tile scans over all nodes would dominate,
    // the stable radix sort of the generic path is the better builder there)
    return range_list && !edge_type && entries > 0 && n_rel > 1 && n_rel <= 65535 && grp_warps(n_nodes) > 0 &&
           place_warps(n_nodes, n_other) > 0 && (!doubled || n_nodes == n_other) &&
           n_nodes * n_rel <= 4 * entries + (int64_t(1) << 20);
}

size_t csr_build_ws_bytes(int64_t entries, int64_t n_nodes, int64_t n_rel) {
    CsrLayout L = make_layout(entries, n_nodes, n_rel);
    int64_t m = entries > L.seg_cap ? entries : L.seg_cap;
    const int64_t dense = n_nodes * n_rel;
    if (dense <= 4 * entries + (int64_t(1) << 20) && dense > m) m = dense;  // grouped path scans the dense key space
    size_t arrays = 5 * (((m + 2) * 4 + 255) & ~size_t(255));
    return arrays + sort_ws_bytes(m) + scan_ws_bytes(m + 2) + size_t(n_rel + 64) * 4 + 2048;
}

int csr_build(const int64_t* edge_index, const int64_t* edge_type, const int64_t* range_list, int64_t E,
              int64_t n_nodes, int64_t n_other, int64_t n_rel, int by_src, int doubled, int drop_loops, int rel_major,
              void* plan, void* ws, cudaStream_t s) {
    const int64_t entries = doubled ? 2 * E : E;
    CsrView v = csr_view(plan, entries, n_nodes, n_rel);
    const uint32_t sentinel = uint32_t(n_nodes * n_rel);
    const int T = 256;
    int64_t m = entries > v.seg_cap ? entries : v.seg_cap;
    const int64_t dense = n_nodes * n_rel;
    if (dense <= 4 * entries + (int64_t(1) << 20) && dense > m) m = dense;  // same rule as csr_build_ws_bytes
    Carver c(ws);
    uint32_t* k0 = c.take<uint32_t>(m + 2);
    uint32_t* v0 = c.take<uint32_t>(m + 2);
    uint32_t* k1 = c.take<uint32_t>(m + 2);
    uint32_t* v1 = c.take<uint32_t>(m + 2);
    int* flags = c.take<int>(m + 2);
    void* sort_ws = c.take<char>(sort_ws_bytes(m));
    void* scan_ws = c.take<char>(scan_ws_bytes(m + 2));
    int* order = c.take<int>(n_rel + 1);

    TIPB_CHECK_CUDA(cudaMemsetAsync(v.counts, 0, 16 * sizeof(int), s));
    TIPB_CHECK_CUDA(cudaMemsetAsync(v.seg_ptr, 0, (v.seg_cap + 1) * sizeof(int), s));
    int rc;
    if (use_grouped(edge_type, range_list, entries, n_nodes, n_other, n_rel, doubled)) {
        const int64_t n_keys = n_nodes * n_rel;
        int* cnt = reinterpret_cast<int*>(k0);
        int* seg_start = reinterpret_cast<int*>(v0);
        int* cnt_fwd = reinterpret_cast<int*>(k1);
        int* seg_index = flags;
        TIPB_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (n_keys + 1) * sizeof(int), s));
        if (doubled) TIPB_CHECK_CUDA(cudaMemsetAsync(cnt_fwd, 0, (n_keys + 1) * sizeof(int), s));
        k_grp_check_ranges<<<(unsigned)ceil_div(n_rel, T), T, 0, s>>>(range_list, (int)n_rel, E, v.counts);
        {
            const size_t hsm = size_t(2) * n_nodes * sizeof(int);
            if ((rc = ensure_dyn_smem((const void*)k_grp_count, hsm))) return rc;
            k_grp_count<<<dim3(4, (unsigned)n_rel), 256, hsm, s>>>(edge_index, range_list, E, doubled, (int)n_nodes,
                                                                  (int)n_other, (int)n_rel, by_src, drop_loops,
                                                                  rel_major, cnt, cnt_fwd, v.counts);
        }
        if ((rc = exclusive_scan_i32(cnt, seg_start, n_keys, scan_ws, s))) return rc;
        const int n_a = int(rel_major ? n_rel : n_nodes), n_b = int(rel_major ? n_nodes : n_rel);
        int* pos_t = reinterpret_cast<int*>(v1);
        k_grp_flags<<<(unsigned)ceil_div(n_keys, T), T, 0, s>>>(cnt, n_keys, n_a, n_b, seg_index, pos_t);
        if ((rc = exclusive_scan_i32(seg_index, seg_index, n_keys, scan_ws, s))) return rc;
        if ((rc = exclusive_scan_i32(pos_t, pos_t, n_keys, scan_ws, s))) return rc;
        k_grp_segments<<<(unsigned)ceil_div(n_keys + 1, T), T, 0, s>>>(cnt, seg_start, seg_index, pos_t, n_keys,
                                                                       (int)n_nodes, (int)n_rel, rel_major, v.seg_ptr,
                                                                       v.seg_node, v.seg_rel, v.rel_seg, v.counts);
        {
            int64_t span = v.seg_cap > n_a + 1 ? v.seg_cap : n_a + 1;
            if (n_b + 1 > span) span = n_b + 1;
            k_grp_tail<<<(unsigned)ceil_div(span, T), T, 0, s>>>(seg_index, pos_t, v.seg_cap, n_a, n_b, (int)n_nodes,
                                                                 (int)n_rel, rel_major ? v.rel_seg_ptr : v.node_ptr,
                                                                 rel_major ? v.node_ptr : v.rel_seg_ptr, v.seg_node,
                                                                 v.seg_rel, v.rel_seg);
        }
        const int pw = place_warps(n_nodes, n_other);
        if (pw) {
            const size_t psm = place_smem(pw, n_nodes);
            const int kb = bits_for(uint64_t(n_nodes - 1));
            k_rel_order<<<(unsigned)ceil_div(n_rel, T), T, 0, s>>>(range_list, (int)n_rel, order);
#define GRP_PLACE(WV)                                                                                               \
            {                                                                                                       \
                auto kern = k_grp_place<WV>;                                                                        \
                if ((rc = ensure_dyn_smem((const void*)kern, psm))) return rc;                                      \
                kern<<<(unsigned)(n_rel * (doubled ? 2 : 1)), WV * 32, psm, s>>>(                                   \
                    edge_index, range_list, order, E, (int)n_nodes, (int)n_other, (int)n_rel, by_src, drop_loops,   \
                    rel_major, doubled, kb, seg_start, cnt_fwd, v.eid, v.other);                                    \
            }
            if (pw == 8) GRP_PLACE(8) else GRP_PLACE(4)
#undef GRP_PLACE
        }
        const int warps = pw ? 0 : grp_warps(n_nodes);
        const size_t smem = size_t(warps + 1) * n_nodes * sizeof(int);
#define GRP_SCATTER(WV)                                                                                             \
        {                                                                                                           \
            auto kern = k_grp_scatter<WV>;                                                                          \
            if ((rc = ensure_dyn_smem((const void*)kern, smem))) return rc;                                         \
            kern<<<dim3((unsigned)n_rel, doubled ? 2 : 1), WV * 32, smem, s>>>(                                      \
                edge_index, range_list, E, (int)n_nodes, (int)n_other, (int)n_rel, by_src, drop_loops, rel_major,   \
                seg_start, cnt_fwd, v.eid, v.other);                                                                \
        }
        if (pw) {} else if (warps == 8) GRP_SCATTER(8) else if (warps == 4) GRP_SCATTER(4) else GRP_SCATTER(2)
#undef GRP_SCATTER
        const unsigned gn = (unsigned)ceil_div(n_nodes, T);
        if (rel_major)
            k_csr_degrees_listed<<<(unsigned)ceil_div(n_nodes * 32, T), T, 0, s>>>((int)n_nodes, v.node_ptr, v.rel_seg, v.seg_ptr, v.deg, v.inv_deg);
        else
            k_csr_degrees<<<gn, T, 0, s>>>((int)n_nodes, v.node_ptr, v.seg_ptr, v.deg, v.inv_deg);
        TIPB_CHECK_LAUNCH("typed_csr_build (grouped)");
        return TIPB_OK;
    } else if (entries > 0) {
        unsigned g = (unsigned)ceil_div(entries, T);
        k_csr_keys<<<g, T, 0, s>>>(edge_index, edge_type, range_list, E, entries, (int)n_nodes, (int)n_other,
                                    (int)n_rel, by_src, drop_loops, rel_major, k0, v0, v.counts);
        if ((rc = sort_pairs_u32(k0, v0, k1, v1, entries, bits_for(sentinel), sort_ws, s))) return rc;
        k_copy_u32_to_i32<<<g, T, 0, s>>>(v1, v.eid, entries);
        k_csr_flags<<<g, T, 0, s>>>(edge_index, E, entries, by_src, sentinel, k1, v.eid, v.other, flags);
        if ((rc = exclusive_scan_i32(flags, flags, entries, scan_ws, s))) return rc;
        k_csr_segments<<<(unsigned)ceil_div(entries + 1, T), T, 0, s>>>(entries, sentinel, (int)n_nodes, (int)n_rel,
                                                                        rel_major, k1, flags, v.seg_ptr, v.seg_node,
                                                                        v.seg_rel, v.counts);
    }
    // primary group pointers (direct segment ranges) and the secondary listing (segment ids grouped the other way)
    unsigned gs = (unsigned)ceil_div(v.seg_cap, T);
    unsigned gp = (unsigned)ceil_div(v.seg_cap + 1, T);
    k_csr_pad<<<gs, T, 0, s>>>(v.seg_cap, (int)n_nodes, (int)n_rel, rel_major, v.counts, v.seg_node, v.seg_rel, k0, v0);
    if (rel_major) {
        k_group_ptr<int><<<gp, T, 0, s>>>(v.seg_rel, v.seg_cap, (int)n_rel, v.rel_seg_ptr);
        if ((rc = sort_pairs_u32(k0, v0, k1, v1, v.seg_cap, bits_for((uint64_t)n_nodes), sort_ws, s))) return rc;
        k_copy_u32_to_i32<<<gs, T, 0, s>>>(v1, v.rel_seg, v.seg_cap);
        k_group_ptr<uint32_t><<<gp, T, 0, s>>>(k1, v.seg_cap, (int)n_nodes, v.node_ptr);
        k_csr_degrees_listed<<<(unsigned)ceil_div(n_nodes * 32, T), T, 0, s>>>((int)n_nodes, v.node_ptr, v.rel_seg, v.seg_ptr,
                                                                          v.deg, v.inv_deg);
    } else {
        k_group_ptr<int><<<gp, T, 0, s>>>(v.seg_node, v.seg_cap, (int)n_nodes, v.node_ptr);
        k_csr_degrees<<<(unsigned)ceil_div(n_nodes, T), T, 0, s>>>((int)n_nodes, v.node_ptr, v.seg_ptr, v.deg, v.inv_deg);
        if ((rc = sort_pairs_u32(k0, v0, k1, v1, v.seg_cap, bits_for((uint64_t)n_rel), sort_ws, s))) return rc;
        k_copy_u32_to_i32<<<gs, T, 0, s>>>(v1, v.rel_seg, v.seg_cap);
        k_group_ptr<uint32_t><<<gp, T, 0, s>>>(k1, v.seg_cap, (int)n_rel, v.rel_seg_ptr);
    }
    TIPB_CHECK_LAUNCH("typed_csr_build");
    return TIPB_OK;
}

}  // namespace tipb

extern "C" {

size_t tipb_typed_csr_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel) {
    return (size_t)tipb::make_layout(n_entries, n_nodes, n_rel).total_bytes;
}

size_t tipb_typed_csr_workspace_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel) {
    return tipb::csr_build_ws_bytes(n_entries, n_nodes, n_rel);
}

int tipb_typed_csr_layout(int64_t n_entries, int64_t n_nodes, int64_t n_rel, int64_t* offsets_bytes,
                          int64_t* seg_capacity) {
    tipb::CsrLayout L = tipb::make_layout(n_entries, n_nodes, n_rel);
    for (int f = 0; f < TIPB_CSR_NFIELDS; ++f) offsets_bytes[f] = L.off[f];
    if (seg_capacity) *seg_capacity = L.seg_cap;
    return TIPB_OK;
}

int tipb_typed_csr_build(const int64_t* edge_index, const int64_t* edge_type, const int64_t* range_list,
                         int64_t n_edges, int64_t n_nodes, int64_t n_other, int64_t n_rel, int by_src, int doubled,
                         int drop_self_loops, int rel_major, void* plan, size_t plan_bytes, void* ws, size_t ws_bytes,
                         void* stream) {
    TIPB_CHECK_ARG(n_edges >= 0 && n_nodes > 0 && n_other > 0 && n_rel > 0, "typed_csr_build: bad sizes");
    const int64_t entries = doubled ? 2 * n_edges : n_edges;
    TIPB_CHECK_ARG(entries < (int64_t(1) << 31) - 2, "typed_csr_build: too many entries for int32 indexing");
    TIPB_CHECK_ARG(n_nodes * n_rel < (int64_t(1) << 32) - 1, "typed_csr_build: n_nodes*n_rel must fit 32 bits");
    TIPB_CHECK_ARG(n_edges == 0 || edge_index, "typed_csr_build: edge_index is NULL");
    TIPB_CHECK_ARG(n_rel == 1 || n_edges == 0 || edge_type || range_list, "typed_csr_build: need edge_type or range_list");
    TIPB_CHECK_ARG(plan && plan_bytes >= tipb_typed_csr_bytes(entries, n_nodes, n_rel), "typed_csr_build: plan buffer too small");
    TIPB_CHECK_ARG(ws && ws_bytes >= tipb_typed_csr_workspace_bytes(entries, n_nodes, n_rel), "typed_csr_build: workspace too small");
    return tipb::csr_build(edge_index, edge_type, range_list, n_edges, n_nodes, n_other, n_rel, by_src, doubled,
                           drop_self_loops, rel_major != 0, plan, ws, (cudaStream_t)stream);
}
}
