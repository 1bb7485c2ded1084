// Fused decoder / loss / gradient pass over PAIRS (north_star item 4: fused gather-bilinear-sigmoid-BCE), for drug
// graphs whose embedding matrix fits in shared memory (the polypharmacy shape: 645 x 16 fp32 = 41 KB).
//
// Replaces, per training step (reference src/layers.py:333-340 and what autograd derives from it):
//   * the typed CSR of the freshly sampled negatives (16.6 M pair-ends counted, scanned and placed by ~15 launches),
//   * both k_decoder_seg passes (one warp per (node, relation) segment: 9-14 warp instructions per pair-end, most of
//     them shuffles that transpose per-entry scalars across the lanes that share a row),
//   * their node-major / relation-major reductions.
//
// A pair (i, j) of relation r is scored ONCE, by one thread:  v = sum_k z[i,k] w[r,k] z[j,k]  (rows from shared
// memory, no shuffles), s = sigmoid(v), the loss term and g = d(loss)/dv.  Its gradient lands on BOTH endpoints:
//   d_z[i] += g w_r * z[j],   d_z[j] += g w_r * z[i],   d_w[r] += g z[i] * z[j].
// To keep that scatter free of floating-point atomics the 2T pair-ends of a tile of T pairs are grouped by node
// inside the CTA (stable counting sort: per-warp ballot ranking, prefix over warps, scan over nodes -- input order
// inside a group, so every sum has ONE fixed order), and each (node, float4-column) accumulator is owned by one
// thread for the whole work item: the Q = dim/4 lanes of a lane group share a node (one float4 column each), so a group's
// row gather is one contiguous 16*Q-byte read, the index / gradient loads are broadcasts inside the group, and a warp's
// trip count is the longest list among 32/Q nodes, not among 32.  A work item is a chunk of <= PAIR_CHUNK pairs of ONE relation (host-built table,
// largest first); it writes  wacc[slot] = w_r * acc  (N x dim),  zacc[slot] = sum_n z[n] * acc[n]  (dim) and its loss
// partial; slots are numbered relation-major, so the reductions that follow are plain fixed-order sums.
//
// Positive edges in the reference layout (src/utils.py:17-23: every relation range is [pairs..., mirrored pairs...])
// are scored through the same kernel: the first half of each range lists every undirected pair once, weight 2.
#include "common.cuh"

namespace tipb {

constexpr float PAIR_EPS = 1e-13f;       // reference src/layers.py:15
constexpr int PP_THREADS = 512;
constexpr int PP_WARPS = PP_THREADS / 32;
constexpr int PP_T = 2048;               // pairs per tile
constexpr int PP_E = 2 * PP_T;           // pair-ends per tile: [0, T) row side, [T, 2T) column side
constexpr int PP_ROUNDS = PP_E / PP_THREADS;   // pair-ends (and ranking rounds) per thread / warp: 8

// row `row`, float4 column q of a [n][Q] float4 matrix staged with an XOR swizzle on q: rows are 16*Q bytes apart, so
// without it the 8 lanes of a quarter-warp gathering random rows at one q would share 2 (Q = 4) bank groups
template <int Q>
__device__ __forceinline__ int zswz(int row, int q) {
    return row * Q + (Q >= 4 ? (q ^ ((row >> 1) & 3)) : q);
}

// one ranking round: among the 32 lanes of a warp, `peers` = lanes holding the same KB-bit key (valid lanes only)
template <int KB>
__device__ __forceinline__ unsigned same_key_lanes(int key, bool valid) {
    unsigned peers = __ballot_sync(FULL, valid);
#pragma unroll
    for (int b = 0; b < KB; ++b) {
        const unsigned bit = (unsigned(key) >> b) & 1u;
        const unsigned m = __ballot_sync(FULL, bit != 0u);
        peers &= ~(m ^ (0u - bit));            // keep the lanes whose bit b equals mine
    }
    return peers;
}

constexpr int PP_NEG_FLAG = 1 << 30;      // items[].w = slot | PP_NEG_FLAG for a work item of the negative pass

template <int DIM, int NPT, int KB>
__global__ void __launch_bounds__(PP_THREADS, DIM >= 32 ? 1 : 2)
k_pair_pass(const uint32_t* __restrict__ pos_pairs, const uint32_t* __restrict__ neg_pairs,
            const int4* __restrict__ items, int n_nodes, const float4* __restrict__ z, const float4* __restrict__ w,
            float pos_weight, float neg_weight, float4* __restrict__ wacc, float4* __restrict__ zacc,
            float* __restrict__ loss_part) {
    constexpr int Q = DIM / 4;
    extern __shared__ float4 pp_smem[];
    float4* zs = pp_smem;                                              // [n_nodes * Q], swizzled
    float4* acc2 = zs + size_t(n_nodes) * Q;                           // [(n_nodes - THREADS) * Q] rows of the nodes >= THREADS
    const int n_hi = NPT > 1 ? max(n_nodes - PP_THREADS, 0) : 0;
    uint32_t* pk = reinterpret_cast<uint32_t*>(acc2 + size_t(n_hi) * Q);    // [T] packed (row << 16 | col)
    float* gb = reinterpret_cast<float*>(pk + PP_T);                   // [T] d(loss)/d(value) of the pair
    uint32_t* lists = reinterpret_cast<uint32_t*>(gb + PP_T);          // [2T] (other << 16 | t), grouped by node
    int* lstart = reinterpret_cast<int*>(lists + PP_E);                // [n_nodes + 1] group offsets of this tile
    uint16_t* wh = reinterpret_cast<uint16_t*>(lstart + ((n_nodes + 1 + 3) & ~3));   // [WARPS][n_nodes], 16-byte aligned
    __shared__ int s_scan[PP_WARPS + 1];
    __shared__ float s_loss[PP_WARPS];
    __shared__ float4 s_zacc[PP_WARPS][Q];
    __shared__ float4 s_w[Q];                  // w[rel]: read as a broadcast

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int4 item = items[blockIdx.x];       // rel, first pair, pair count, slot | pass flag
    const int rel = item.x, p0 = item.y, np = item.z, slot = item.w & (PP_NEG_FLAG - 1);
    const bool negative = (item.w & PP_NEG_FLAG) != 0;          // uniform over the CTA
    const uint32_t* __restrict__ pairs = (negative ? neg_pairs : pos_pairs) + size_t(p0);
    const float pair_weight = negative ? neg_weight : pos_weight;

    for (int i = tid; i < n_nodes * Q; i += PP_THREADS) zs[zswz<Q>(i / Q, i % Q)] = z[i];
    for (int i = tid; i < n_hi * Q; i += PP_THREADS) acc2[i] = f4_zero();
    if (tid < Q) s_w[tid] = w[size_t(rel) * Q + tid];

    // accumulator units: unit u = k * THREADS + tid is (node u / Q, float4 column u % Q = tid % Q); a thread owns up to
    // Q * NPT units: the first Q in registers, the rest (nodes >= THREADS) in shared memory (acc2, same unit order)
    constexpr int GROUP_NODES = PP_THREADS / Q;          // nodes covered by one round of units
    const int myq = tid % Q, mynode0 = tid / Q;
    float4 acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = f4_zero();
    float loss = 0.f;
    uint16_t* mywh = wh + wid * n_nodes;
    const int per = (n_nodes + PP_THREADS - 1) / PP_THREADS;   // nodes per thread in the scan
    const int wh_vec = (PP_WARPS * n_nodes * 2 + 15) / 16;     // uint4 words of wh

    // the next tile's pairs are fetched into registers while the current tile is processed
    uint32_t nxt[PP_T / PP_THREADS];
#pragma unroll
    for (int u = 0; u < PP_T / PP_THREADS; ++u) {
        const int t = tid + u * PP_THREADS;
        nxt[u] = t < np ? ld_stream_i32(reinterpret_cast<const int*>(pairs) + t) : 0u;
    }
    for (int t0 = 0; t0 < np; t0 += PP_T) {
        const int nt = min(PP_T, np - t0);
        __syncthreads();                       // previous tile fully consumed (and zs staged, first trip)
#pragma unroll
        for (int u = 0; u < PP_T / PP_THREADS; ++u) {
            const int t = tid + u * PP_THREADS;
            pk[t] = nxt[u];
            const int tn = t0 + PP_T + t;
            nxt[u] = tn < np ? ld_stream_i32(reinterpret_cast<const int*>(pairs) + tn) : 0u;
        }
        for (int i = tid; i < wh_vec; i += PP_THREADS) reinterpret_cast<uint4*>(wh)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();

        // ---- phase 1: one thread scores one pair (all operands in shared memory, no shuffles)
        for (int t = tid; t < nt; t += PP_THREADS) {
            const uint32_t p = pk[t];
            const int i = int(p >> 16), j = int(p & 0xffffu);
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) v += f4_dot(f4_mul(zs[zswz<Q>(i, q)], s_w[q]), zs[zswz<Q>(j, q)]);
            const float sg = __frcp_rn(1.0f + __expf(-v));
            // positives: -log(s + eps), d/dv = -s (1 - s) / (s + eps);  negatives: -log(1 - s + eps), d/dv = s (1 - s) / (1 - s + eps)
            const float om = 1.f - sg;
            const float arg = (negative ? om : sg) + PAIR_EPS;
            loss -= __logf(arg);
            const float g = __fdividef(sg * om, arg);
            gb[t] = negative ? g * pair_weight : -g * pair_weight;
        }

        // ---- group the 2 nt pair-ends by node: count (stable rank inside the warp's run) ...
        // warp w owns pair-ends [w * 256, (w + 1) * 256) in 8 rounds of 32; (warp, round, lane) order is input order
        uint32_t rk8[PP_ROUNDS / 4] = {0u, 0u};  // rank inside the warp's 256-entry run (< 256): one byte per round
        static_assert(PP_E / PP_WARPS <= 256 && PP_ROUNDS == 8, "ranks are packed as bytes");
#pragma unroll
        for (int rd = 0; rd < PP_ROUNDS; ++rd) {
            const int e = wid * (PP_E / PP_WARPS) + rd * 32 + lane;
            const int t = e & (PP_T - 1);
            const bool valid = t < nt;
            const uint32_t p = pk[valid ? t : 0];
            const int node = int(e < PP_T ? (p >> 16) : (p & 0xffffu));
            const unsigned peers = same_key_lanes<KB>(node, valid);
            int c0 = 0;
            if (valid) c0 = mywh[node];
            __syncwarp();
            const int rank = __popc(peers & lt);
            if (valid && rank == 0) mywh[node] = uint16_t(c0 + __popc(peers));
            __syncwarp();
            rk8[rd >> 2] |= uint32_t(c0 + rank) << ((rd & 3) * 8);
        }
        __syncthreads();
        // ... scan: per node a prefix over the warps, then an exclusive scan over the nodes
        int mine = 0;
        for (int k = 0; k < per; ++k) {
            const int n = tid * per + k;
            if (n < n_nodes) {
                int run = 0;
#pragma unroll
                for (int ww = 0; ww < PP_WARPS; ++ww) {
                    const int c = wh[ww * n_nodes + n];
                    wh[ww * n_nodes + n] = uint16_t(run);
                    run += c;
                }
                lstart[n] = run;               // the node's count for now
                mine += run;
            }
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_scan[wid] = incl;
        __syncthreads();
        if (wid == 0) {                        // exclusive scan of the 16 warp totals by one warp
            const int c = lane < PP_WARPS ? s_scan[lane] : 0;
            int x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, x, o);
                if (lane >= o) x += y;
            }
            if (lane < PP_WARPS) s_scan[lane] = x - c;
            if (lane == PP_WARPS - 1) s_scan[PP_WARPS] = x;
        }
        __syncthreads();
        int ls = s_scan[wid] + incl - mine;
        for (int k = 0; k < per; ++k) {
            const int n = tid * per + k;
            if (n < n_nodes) {
                const int c = lstart[n];
                lstart[n] = ls;
                ls += c;
            }
        }
        if (tid == 0) lstart[n_nodes] = s_scan[PP_WARPS];
        __syncthreads();
        // ... place
#pragma unroll
        for (int rd = 0; rd < PP_ROUNDS; ++rd) {
            const int e = wid * (PP_E / PP_WARPS) + rd * 32 + lane;
            const int t = e & (PP_T - 1);
            if (t < nt) {
                const uint32_t p = pk[t];
                const uint32_t hi = p >> 16, lo = p & 0xffffu;
                const int node = int(e < PP_T ? hi : lo);
                const uint32_t other = e < PP_T ? lo : hi;
                const int rank = int((rk8[rd >> 2] >> ((rd & 3) * 8)) & 0xffu);
                lists[lstart[node] + mywh[node] + rank] = (other << 16) | uint32_t(t);
            }
        }
        __syncthreads();

        // ---- phase 2: every (node, column) unit adds its node's group, in list order.  The Q lanes of a group read the
        // same list entry / gradient (broadcast) and the Q consecutive float4 of one row (one contiguous read)
#pragma unroll
        for (int k = 0; k < Q * NPT; ++k) {
            const int n = k * GROUP_NODES + mynode0;
            if (n < n_nodes) {
                float4 a = k < Q ? acc[k < Q ? k : 0] : acc2[(k - Q) * PP_THREADS + tid];
                const int end = lstart[n + 1];
                for (int idx = lstart[n]; idx < end; ++idx) {
                    const uint32_t en = lists[idx];
                    a = f4_fma(gb[en & 0xffffu], zs[zswz<Q>(int(en >> 16), myq)], a);
                }
                if (k < Q) acc[k < Q ? k : 0] = a; else acc2[(k - Q) * PP_THREADS + tid] = a;
            }
        }
    }

    // ---- item epilogue: wacc[slot] = w_r * acc, zacc[slot] = sum_n z[n] * acc[n], loss partial (fixed orders)
    float4 zsum = f4_zero();                   // this thread's column (myq) of sum_n z[n] * acc[n]
    float4* wout = wacc + size_t(slot) * n_nodes * Q;
#pragma unroll
    for (int k = 0; k < Q * NPT; ++k) {        // (this thread's own units: no barrier needed)
        const int n = k * GROUP_NODES + mynode0;
        if (n < n_nodes) {
            const float4 a = k < Q ? acc[k < Q ? k : 0] : acc2[(k - Q) * PP_THREADS + tid];
            wout[n * Q + myq] = f4_mul(a, s_w[myq]);
            zsum = f4_add(zsum, f4_mul(a, zs[zswz<Q>(n, myq)]));
        }
    }
#pragma unroll
    for (int o = 16; o >= Q; o >>= 1) {        // lanes with the same lane % Q hold the same column
        zsum.x += __shfl_xor_sync(FULL, zsum.x, o);
        zsum.y += __shfl_xor_sync(FULL, zsum.y, o);
        zsum.z += __shfl_xor_sync(FULL, zsum.z, o);
        zsum.w += __shfl_xor_sync(FULL, zsum.w, o);
    }
    loss = warp_sum(loss);
    if (lane < Q) s_zacc[wid][lane] = zsum;    // lane % Q == lane here
    if (lane == 0) s_loss[wid] = loss;
    __syncthreads();
    if (tid < Q) {
        float4 t = s_zacc[0][tid];
#pragma unroll
        for (int ww = 1; ww < PP_WARPS; ++ww) t = f4_add(t, s_zacc[ww][tid]);
        zacc[size_t(slot) * Q + tid] = t;
    }
    if (tid == 32) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < PP_WARPS; ++ww) t += s_loss[ww];
        loss_part[slot] = t * pair_weight;
    }
}

// d_z partial sums: part[g][c] = sum over the slots of group g of wacc[slot][c]   (c < n_cells = N * dim / 4 float4)
__global__ void __launch_bounds__(256)
k_pair_reduce_dz(const float4* __restrict__ wacc, int n_slots, int n_cells, int n_groups, float4* __restrict__ part) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n_cells) return;
    const int g = blockIdx.y;
    const int per = (n_slots + n_groups - 1) / n_groups;
    const int s0 = g * per, s1 = min(n_slots, s0 + per);
    float4 a0 = f4_zero(), a1 = f4_zero(), a2 = f4_zero(), a3 = f4_zero();
    int s = s0;
    for (; s + 3 < s1; s += 4) {
        a0 = f4_add(a0, wacc[size_t(s) * n_cells + c]);
        a1 = f4_add(a1, wacc[size_t(s + 1) * n_cells + c]);
        a2 = f4_add(a2, wacc[size_t(s + 2) * n_cells + c]);
        a3 = f4_add(a3, wacc[size_t(s + 3) * n_cells + c]);
    }
    for (; s < s1; ++s) a0 = f4_add(a0, wacc[size_t(s) * n_cells + c]);
    part[size_t(g) * n_cells + c] = f4_add(f4_add(a0, a1), f4_add(a2, a3));
}

// d_z = sum_g part[g];  d_w[r] = 1/2 sum_{slots of r} zacc[slot];  loss = sum_slots loss_part   (one launch)
__global__ void __launch_bounds__(256)
k_pair_finish(const float4* __restrict__ part, int n_groups, int n_cells, const float* __restrict__ zacc,
              const int* __restrict__ rel_slot_ptr, int n_rel, int dim, const float* __restrict__ loss_part, int n_slots,
              float4* __restrict__ d_z, float* __restrict__ d_w, float* __restrict__ loss_out) {
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid < n_cells) {
        float4 a = part[gid];
        for (int g = 1; g < n_groups; ++g) a = f4_add(a, part[size_t(g) * n_cells + gid]);
        d_z[gid] = a;
    }
    if (gid < n_rel * dim) {
        const int r = gid / dim, k = gid % dim;
        float a = 0.f;
        for (int s = rel_slot_ptr[r]; s < rel_slot_ptr[r + 1]; ++s) a += zacc[size_t(s) * dim + k];
        // both passes number their slots relation-major one after the other: [pos slots of all r | neg slots of all r]
        for (int s = rel_slot_ptr[n_rel + 1 + r]; s < rel_slot_ptr[n_rel + 1 + r + 1]; ++s) a += zacc[size_t(s) * dim + k];
        d_w[gid] = 0.5f * a;
    }
    if (blockIdx.x == 0) {
        __shared__ float sw[8];
        float a = 0.f;
        for (int s = threadIdx.x; s < n_slots; s += 256) a += loss_part[s];
        a = warp_sum(a);
        if (lane_id() == 0) sw[warp_id()] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < 8; ++k) t += sw[k];
            loss_out[0] = t;
        }
    }
}

// first half of every relation range of a mirrored edge set -> packed pairs (row << 16 | col); pair e/2-ordered
__global__ void __launch_bounds__(256)
k_pack_half_pairs(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ range_list, int64_t n_edges,
                  int n_rel, int n_nodes, uint32_t* __restrict__ out, int* __restrict__ status) {
    const int r = blockIdx.y;
    const int64_t s = range_list[2 * r], e = range_list[2 * r + 1];
    const int64_t half = (e - s) >> 1;
    for (int64_t x = int64_t(blockIdx.x) * 256 + threadIdx.x; x < half; x += int64_t(gridDim.x) * 256) {
        const int64_t a = edge_index[s + x], b = edge_index[n_edges + s + x];
        if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) { atomicOr(status, 1); continue; }
        out[(s >> 1) + x] = (uint32_t(a) << 16) | uint32_t(b);
    }
}

// packed pairs -> int64 [2, n] (the reference's LongTensor layout), for callers that want to look at the sample
__global__ void k_unpack_pairs(const uint32_t* __restrict__ packed, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = packed[i];
    out[i] = int64_t(p >> 16);
    out[n + i] = int64_t(p & 0xffffu);
}

static size_t pair_smem_bytes(int64_t n_nodes, int dim) {
    const int64_t n_hi = n_nodes > PP_THREADS ? n_nodes - PP_THREADS : 0;
    size_t b = size_t(n_nodes + n_hi) * dim * 4;       // zs, acc2
    b += size_t(PP_T) * 4 * 2 + size_t(PP_E) * 4;      // pk, gb, lists
    b += size_t((n_nodes + 1 + 3) & ~int64_t(3)) * 4;  // lstart
    b += (size_t(PP_WARPS) * n_nodes * 2 + 15) & ~size_t(15);   // wh
    return (b + 15) & ~size_t(15);
}

constexpr int PAIR_DZ_GROUPS = 24;

static bool pair_shape_ok(int64_t n_nodes, int dim) {
    if (!(dim == 4 || dim == 8 || dim == 16 || dim == 32)) return false;
    if (n_nodes < 1 || n_nodes > 2 * PP_THREADS) return false;                  // <= 2 accumulator rows per thread
    return pair_smem_bytes(n_nodes, dim) + 4096 <= size_t(max_smem_optin());   // + static shared memory, 1 CTA per SM
}

template <int DIM, int NPT, int KB>
static int pair_pass_launch(const uint32_t* pos_pairs, const uint32_t* neg_pairs, const int4* items, int n_items,
                            int n_nodes, const float* z, const float* w, float pos_weight, float neg_weight, float* wacc,
                            float* zacc, float* loss_part, cudaStream_t s) {
    auto kern = k_pair_pass<DIM, NPT, KB>;
    const size_t smem = pair_smem_bytes(n_nodes, DIM);
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    kern<<<n_items, PP_THREADS, smem, s>>>(pos_pairs, neg_pairs, items, n_nodes, (const float4*)z, (const float4*)w,
                                           pos_weight, neg_weight, (float4*)wacc, (float4*)zacc, loss_part);
    TIPB_CHECK_LAUNCH("pair_pass");
    return TIPB_OK;
}

template <int DIM>
static int pair_pass_pick(const uint32_t* pos_pairs, const uint32_t* neg_pairs, const int4* items, int n_items,
                          int n_nodes, const float* z, const float* w, float pos_weight, float neg_weight, float* wacc,
                          float* zacc, float* loss_part, cudaStream_t s) {
#define PP_GO(N, K) return pair_pass_launch<DIM, N, K>(pos_pairs, neg_pairs, items, n_items, n_nodes, z, w, pos_weight, neg_weight, wacc, zacc, loss_part, s)
    if (n_nodes <= 128) PP_GO(1, 7);
    if (n_nodes <= PP_THREADS) PP_GO(1, 9);
    if (n_nodes <= 1024) PP_GO(2, 10);
#undef PP_GO
    set_last_error("pair_pass: n_nodes=%d not supported", n_nodes);
    return TIPB_ERR_UNSUPPORTED;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_pair_pass_supported(int64_t n_nodes, int dim) { return pair_shape_ok(n_nodes, dim) ? 1 : 0; }

int64_t tipb_pair_chunk(void) { return 8 * PP_T; }

size_t tipb_pair_workspace_bytes(int64_t n_slots, int64_t n_nodes, int dim) {
    const size_t cells = size_t(n_nodes) * dim;
    return (size_t(n_slots) * cells + size_t(n_slots) * dim + size_t(n_slots) + size_t(PAIR_DZ_GROUPS) * cells) * 4 + 2048;
}

int tipb_pack_half_pairs(const int64_t* edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_rel,
                         int64_t n_nodes, uint32_t* packed, int32_t* status, void* stream) {
    TIPB_CHECK_ARG(range_list && packed && status && (n_edges == 0 || edge_index), "pack_half_pairs: NULL argument");
    TIPB_CHECK_ARG(n_nodes > 0 && n_nodes <= 65535 && n_rel > 0 && n_rel <= 65535, "pack_half_pairs: sizes out of range");
    k_pack_half_pairs<<<dim3(8, (unsigned)n_rel), 256, 0, (cudaStream_t)stream>>>(edge_index, range_list, n_edges, (int)n_rel,
                                                                                (int)n_nodes, packed, status);
    TIPB_CHECK_LAUNCH("pack_half_pairs");
    return TIPB_OK;
}

int tipb_unpack_pairs(const uint32_t* packed, int64_t n, int64_t* edge_index, void* stream) {
    TIPB_CHECK_ARG(n == 0 || (packed && edge_index), "unpack_pairs: NULL argument");
    if (n > 0) k_unpack_pairs<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(packed, n, edge_index);
    TIPB_CHECK_LAUNCH("unpack_pairs");
    return TIPB_OK;
}

// both passes in one launch: work items of the positive pairs and of the negative pairs (flagged in items[].w)
int tipb_pair_bce_pass(const uint32_t* pos_pairs, const uint32_t* neg_pairs, const int32_t* items, int64_t n_items,
                       int64_t n_slots_total, int64_t n_nodes, const float* z, const float* weight, int dim,
                       float pos_weight, float neg_weight, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(pos_pairs && neg_pairs && items && z && weight && ws, "pair_bce_pass: NULL argument");
    TIPB_CHECK_ARG(pair_shape_ok(n_nodes, dim), "pair_bce_pass: shape not supported (use the typed-CSR decoder path)");
    TIPB_CHECK_ARG(n_slots_total < PP_NEG_FLAG, "pair_bce_pass: too many slots");
    TIPB_CHECK_ARG(ws_bytes >= tipb_pair_workspace_bytes(n_slots_total, n_nodes, dim), "pair_bce_pass: workspace too small");
    if (n_items == 0) return TIPB_OK;
    const size_t cells = size_t(n_nodes) * dim;
    float* wacc = static_cast<float*>(ws);
    float* zacc = wacc + size_t(n_slots_total) * cells;
    float* loss_part = zacc + size_t(n_slots_total) * dim;
    cudaStream_t s = (cudaStream_t)stream;
    const int4* it = reinterpret_cast<const int4*>(items);
#define PP_DIM(D)                                                                                                   \
    return pair_pass_pick<D>(pos_pairs, neg_pairs, it, (int)n_items, (int)n_nodes, z, weight, pos_weight, neg_weight, \
                             wacc, zacc, loss_part, s)
    switch (dim) {
        case 4: PP_DIM(4);
        case 8: PP_DIM(8);
        case 16: PP_DIM(16);
        default: PP_DIM(32);
    }
#undef PP_DIM
}

// d_z, d_w, loss from the slots of both passes.  rel_slot_ptr: [2 * (n_rel + 1)] -- the positive pass's relation-major
// slot ranges, then the negative pass's (absolute slot numbers).
int tipb_pair_bce_finish(const int32_t* rel_slot_ptr, int64_t n_slots_total, int64_t n_nodes, int64_t n_rel, int dim,
                         float* loss_out, float* d_z, float* d_weight, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(rel_slot_ptr && loss_out && d_z && d_weight && ws, "pair_bce_finish: NULL argument");
    TIPB_CHECK_ARG(ws_bytes >= tipb_pair_workspace_bytes(n_slots_total, n_nodes, dim), "pair_bce_finish: workspace too small");
    const size_t cells = size_t(n_nodes) * dim;
    float* wacc = static_cast<float*>(ws);
    float* zacc = wacc + size_t(n_slots_total) * cells;
    float* loss_part = zacc + size_t(n_slots_total) * dim;
    float* part = loss_part + n_slots_total;
    part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(part) + 15) & ~uintptr_t(15));
    cudaStream_t s = (cudaStream_t)stream;
    const int n_cells = int(cells / 4);
    int groups = PAIR_DZ_GROUPS;
    if (groups > n_slots_total) groups = int(n_slots_total > 0 ? n_slots_total : 1);
    k_pair_reduce_dz<<<dim3((unsigned)ceil_div(n_cells, 256), groups), 256, 0, s>>>((const float4*)wacc, (int)n_slots_total,
                                                                                  n_cells, groups, (float4*)part);
    int64_t span = n_cells > n_rel * dim ? n_cells : n_rel * dim;
    k_pair_finish<<<(unsigned)ceil_div(span, 256), 256, 0, s>>>((const float4*)part, groups, n_cells, zacc, rel_slot_ptr,
                                                                (int)n_rel, dim, loss_part, (int)n_slots_total,
                                                                (float4*)d_z, d_weight, loss_out);
    TIPB_CHECK_LAUNCH("pair_bce_finish");
    return TIPB_OK;
}
}
