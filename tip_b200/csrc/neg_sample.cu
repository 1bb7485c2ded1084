// Typed negative sampling on the GPU, bit-exact with the reference (north_star item 6, SURVEY.md
// section 8a row N).  Replaces src/neg_sampling.py:5-26, which runs on the CPU through numpy:
//
//   for each relation r (k_r positive pairs, keys row*N+col):
//       perm = np.random.choice(N*N, k_r)
//       rest = positions of perm that are positive pairs
//       while rest not empty:
//           tmp  = np.random.choice(N*N, len(rest));  perm[rest] = tmp
//           rest = positions INSIDE tmp that are positive pairs      (sic: the reference indexes tmp)
//       row = float32(perm) / float32(N) -> trunc,  col = perm % N
//
// np.random.choice(n, k) of the legacy global RandomState is: take successive tempered MT19937
// outputs, AND with (2^ceil(log2 n) - 1), drop values > n-1, until k are kept.  The stream is shared
// by all relations, so relation r starts where relation r-1 (including its retries) stopped.
//
// How this becomes a parallel program
//   1. The raw MT19937 words do not depend on the data: tipb_mt19937_generate produces them ahead of
//      time (the Python host overlaps it with the rest of the training step on a side stream).
//   2. Masked rejection is a stream compaction (flags -> scan -> scatter): accepted stream A.
//   3. A relation's retry loop ends when a whole round has no positive pair, which is exactly when the
//      relation has consumed its k_r-th NON-member value of A.  So the start offset of relation r+1 is
//      o_{r+1} = 1 + position of the k_r-th non-member (w.r.t. relation r's bitmap) at or after o_r.
//      o_r is only known to within the fluctuation of the retry counts, so every relation evaluates that
//      map for a whole bracket of candidate offsets in parallel (k_window_scan; brackets are
//      mean +- z*sigma of a negative-binomial model, built on the host once per graph by
//      tipb_neg_table_build) and one thread then walks the chain of table lookups.
//      A bracket miss is detected and reported (status bit 2); the caller reruns in exact mode
//      (k_chain_exact: one CTA walks the relations and counts round by round).
//   4. With o_r known, round 0 and round 1 of every relation are independent per draw (k_materialize_main),
//      the geometrically smaller rounds >= 2 are replayed by one warp per relation (k_materialize_fixup).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace tipb {

constexpr int MT_N = 624, MT_M = 397, MT_LAG = MT_N - MT_M;  // 227
constexpr uint32_t MT_UPPER = 0x80000000u, MT_LOWER = 0x7fffffffu, MT_MATRIX_A = 0x9908b0dfu;
constexpr int RING = 2048;
constexpr int TAB = 8;  // per-relation table row: lo, W, L, win_off, f_off, k, pred (expected start offset), order
                        // (order: row b holds the relation with the b-th longest window -- launch order of k_window_scan)

enum { NEG_STATUS_OUT_OF_WORDS = 1, NEG_STATUS_TOO_MANY_ROUNDS = 2, NEG_STATUS_BRACKET_MISS = 4 };

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    uint32_t y = (a & MT_UPPER) | (b & MT_LOWER);
    return (y >> 1) ^ (MT_MATRIX_A & (0u - (b & 1u)));
}

__global__ void k_mt_seed(uint32_t* __restrict__ state, uint32_t seed) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t s = seed;
        for (int i = 0; i < MT_N; ++i) {
            state[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + uint32_t(i + 1);
        }
        state[MT_N] = MT_N;
    }
}

// U[0..624) = current key block; U[624 + i] = i-th further word of the untempered stream, i < n_new
// (n_new a multiple of 454).  One CTA runs S[n+624] = S[n+397] ^ mix(S[n], S[n+1]): thread t owns the
// outputs with n = t mod 227, so S[n+397] is its own previous output (a register) and the two mix operands
// were written at least two phases (454 words) earlier -> one barrier per two phases.
__global__ void __launch_bounds__(256) k_mt_generate(const uint32_t* __restrict__ state, uint32_t* __restrict__ U,
                                                     int n_new) {
    __shared__ uint32_t ring[RING];
    const int t = threadIdx.x;
    for (int i = t; i < MT_N; i += 256) {
        const uint32_t v = state[i];
        ring[i] = v;
        U[i] = v;
    }
    __syncthreads();
    const bool active = t < MT_LAG;  // threads 227..255 only take part in the barriers
    uint32_t prev = active ? ring[MT_M + t] : 0u;
    uint32_t* out = U + MT_N;
    const int n_iter = n_new / (2 * MT_LAG);
    int n = active ? t : 0;
    for (int it = 0; it < n_iter; ++it, n += 2 * MT_LAG) {
        if (active) {
            const uint32_t v0 = prev ^ mt_mix(ring[n & (RING - 1)], ring[(n + 1) & (RING - 1)]);
            ring[(n + MT_N) & (RING - 1)] = v0;
            out[n] = v0;
            const int n1 = n + MT_LAG;
            const uint32_t v1 = v0 ^ mt_mix(ring[n1 & (RING - 1)], ring[(n1 + 1) & (RING - 1)]);
            ring[(n1 + MT_N) & (RING - 1)] = v1;
            out[n1] = v1;
            prev = v1;
        }
        __syncthreads();
    }
}

// Masked rejection as a stream compaction.  Word i of the window is a candidate iff it lies at or after the
// state's read position; it is accepted iff (tempered & mask) <= max_val.  Two passes over the raw words, no flag
// array: per-chunk counts -> scan of the (few thousand) chunk counts -> every chunk recomputes its flags and
// writes its accepted values in word order.  Thread t of a chunk owns ACC_ITEMS consecutive words.
constexpr int ACC_THREADS = 256;
constexpr int ACC_ITEMS = 8;
constexpr int ACC_CHUNK = ACC_THREADS * ACC_ITEMS;

__device__ __forceinline__ unsigned accept_bits(const uint32_t* __restrict__ U, int64_t base, int64_t n_words,
                                                int64_t first, uint32_t mask, uint32_t max_val, uint32_t* w) {
    unsigned bits = 0;
    if (base + ACC_ITEMS <= n_words) {
        const uint4 q0 = *reinterpret_cast<const uint4*>(U + base), q1 = *reinterpret_cast<const uint4*>(U + base + 4);
        w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w; w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
    } else {
#pragma unroll
        for (int i = 0; i < ACC_ITEMS; ++i) w[i] = base + i < n_words ? U[base + i] : 0xffffffffu;
    }
#pragma unroll
    for (int i = 0; i < ACC_ITEMS; ++i) {
        w[i] = mt_temper(w[i]) & mask;
        if (base + i < n_words && base + i >= first && w[i] <= max_val) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(ACC_THREADS)
k_accept_count(const uint32_t* __restrict__ U, const uint32_t* __restrict__ pos_ptr, int64_t n_words, uint32_t mask,
               uint32_t max_val, int* __restrict__ chunk_cnt) {
    __shared__ int sw[ACC_THREADS / 32];
    uint32_t w[ACC_ITEMS];
    const int64_t base = int64_t(blockIdx.x) * ACC_CHUNK + int64_t(threadIdx.x) * ACC_ITEMS;
    int c = __popc(accept_bits(U, base, n_words, int64_t(*pos_ptr), mask, max_val, w));
    c = warp_sum_i(c);
    if (lane_id() == 0) sw[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < ACC_THREADS / 32; ++k) t += sw[k];
        chunk_cnt[blockIdx.x] = t;
    }
}

// accepted values of a chunk are compacted in shared memory and written out as whole lines (a thread's own run of up to
// eight values starts at an arbitrary offset: written directly, every store instruction of the warp touched ~25 sectors).
// The word position of an accepted value is needed for ONE value per call only (the last one consumed): k_finalize
// recomputes it from the chunk offsets instead of a second 35 MB array.
__global__ void __launch_bounds__(ACC_THREADS)
k_compact(const uint32_t* __restrict__ U, const uint32_t* __restrict__ pos_ptr, int64_t n_words, uint32_t mask,
          uint32_t max_val, const int* __restrict__ chunk_off, int* __restrict__ A) {
    __shared__ int sw[ACC_THREADS / 32];
    __shared__ int sA[ACC_CHUNK];
    uint32_t w[ACC_ITEMS];
    const int64_t base = int64_t(blockIdx.x) * ACC_CHUNK + int64_t(threadIdx.x) * ACC_ITEMS;
    const unsigned bits = accept_bits(U, base, n_words, int64_t(*pos_ptr), mask, max_val, w);
    const int c = __popc(bits);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, incl, o);
        if (lane_id() >= o) incl += y;
    }
    if (lane_id() == 31) sw[warp_id()] = incl;
    __syncthreads();
    int a = incl - c, total = 0;
#pragma unroll
    for (int k = 0; k < ACC_THREADS / 32; ++k) {
        if (k < warp_id()) a += sw[k];
        total += sw[k];
    }
#pragma unroll
    for (int i = 0; i < ACC_ITEMS; ++i) {
        if ((bits >> i) & 1u) sA[a++] = int(w[i]);
    }
    __syncthreads();
    int* dst = A + chunk_off[blockIdx.x];
    for (int i = threadIdx.x; i < total; i += ACC_THREADS) dst[i] = sA[i];
}

__device__ __forceinline__ bool is_member(const uint32_t* __restrict__ bits, int key) {
    return (bits[key >> 5] >> (key & 31)) & 1u;
}

// `out` = the reference's int64 [2, n_edges] (may be NULL), `packed` = (row << 16 | col) per pair (may be NULL; the
// fused training step consumes this form, csrc/pair_pass.cu)
__device__ __forceinline__ void write_pair(int64_t* __restrict__ out, uint32_t* __restrict__ packed, int64_t n_edges,
                                           int64_t e, int p, int n_nodes, float fn) {
    const float row = __fdiv_rn(__int2float_rn(p), fn);  // float32 true division, as torch does for perm / N
    const long long r = (long long)row;                  // .long(): truncation toward zero
    const int c = p % n_nodes;
    if (out) {
        out[e] = r;
        out[n_edges + e] = (long long)c;
    }
    if (packed) packed[e] = (uint32_t(r) << 16) | uint32_t(c);
}

// ================================================================================================
// fast path
// ================================================================================================
// One CTA per relation.  Window = accepted-stream positions [lo, lo+L).  Produces
//   NHI[win_off + x] = number of non-members among window entries 0..x (inclusive)
//   PR [win_off + m] = window index of the (m+1)-th non-member
//   F  [f_off + x]   = next relation's start offset if this relation starts at lo + x   (x < W; -1: window too short)
// STAGED: the relation's bitmap (n_nodes^2 bits; 52 KB for 645 drugs) is copied into shared memory first.  The probes are
// random 4-byte reads: from global memory each one costs a 32-byte sector and an L1 tag cycle per LANE (the first version
// ran at ~1 probe per clock per SM: 200 us for 15 M window entries), from shared memory a few bank-conflict replays per warp.
//
// Two passes, ONE block barrier between them (the second version scanned 4096 entries per trip with four barriers each):
//   pass 1  every warp owns a contiguous run of 32-entry groups: coalesced loads of A, probe, ballot -> the group's
//           non-member mask (kept in shared memory) and the warp's non-member count;
//   pass 2  base = sum of the earlier warps' counts; per group, nhi and pr follow from the stored mask alone
//           (popc of the lanes at or below / below mine): coalesced stores, no second look at A or the bitmap.
constexpr int WS_THREADS = 1024;
constexpr int WS_WARPS = WS_THREADS / 32;
constexpr int WS_MASK_WORDS = 6144;        // group masks kept in shared memory: windows up to 196,608 entries

template <bool STAGED>
__global__ void __launch_bounds__(WS_THREADS)
k_window_scan(const int* __restrict__ A, const int* __restrict__ n_accepted_ptr, const uint32_t* __restrict__ member,
              int64_t words_per_rel, const int64_t* __restrict__ table, int r_lo, int by_order, int* __restrict__ NHI,
              int* __restrict__ PR, int* __restrict__ F) {
    // `member` holds the bitmaps of the relations [r_lo, ...) only (a rank's shard; r_lo = 0: all relations)
    extern __shared__ uint32_t ws_bits[];
    __shared__ uint32_t s_mask[WS_MASK_WORDS];
    __shared__ int sw[WS_WARPS];
    // unsharded launches (by_order) take the relations longest window first: the windows differ by 40x and the CTAs
    // of a launch start in index order, so a long window met late would run alone at the end
    const int r = by_order ? int(table[int64_t(blockIdx.x) * TAB + 7]) : r_lo + blockIdx.x;
    const int64_t* tb = table + int64_t(r) * TAB;
    const int lo = int(tb[0]), W = int(tb[1]), L = int(tb[2]), k = int(tb[5]);
    const int64_t win_off = tb[3], f_off = tb[4];
    if (k == 0) {
        for (int x = threadIdx.x; x < W; x += WS_THREADS) F[f_off + x] = lo + x;
        return;
    }
    const uint32_t* gbits = member + int64_t(r - r_lo) * words_per_rel;
    const int n_acc = *n_accepted_ptr;
    if (n_acc <= 0) {  // no usable stream at all
        for (int x = threadIdx.x; x < W; x += WS_THREADS) F[f_off + x] = -1;
        return;
    }
    if (STAGED) {
        for (int i = threadIdx.x; i < int(words_per_rel); i += WS_THREADS) ws_bits[i] = gbits[i];
        __syncthreads();
    }
    const uint32_t* bits = STAGED ? ws_bits : gbits;
    int* nhi = NHI + win_off;
    int* pr = PR + win_off;
    const int lane = lane_id(), wid = warp_id();
    const int n_groups = (L + 31) >> 5;
    const int gpw = (n_groups + WS_WARPS - 1) / WS_WARPS;
    const int g0 = min(wid * gpw, n_groups), g1 = min(g0 + gpw, n_groups);
    const bool keep = n_groups <= WS_MASK_WORDS;          // uniform over the CTA
    const int last = n_acc - 1;

    auto group_mask = [&](int g) -> uint32_t {            // non-member mask of group g (all lanes take part)
        const int x = (g << 5) + lane, j = lo + x;
        const int v = A[min(j, last)];
        const bool nm = x < L && j <= last && !((bits[v >> 5] >> (v & 31)) & 1u);
        return __ballot_sync(FULL, nm);
    };

    // ---- pass 1
    int cnt = 0;
    int g = g0;
    for (; g + 4 <= g1; g += 4) {                          // four independent load -> probe chains in flight
        int v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = ((g + u) << 5) + lane, j = lo + x;
            ok[u] = x < L && j <= last;
            v[u] = A[min(j, last)];
        }
        uint32_t wd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wd[u] = bits[v[u] >> 5];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t m = __ballot_sync(FULL, ok[u] && !((wd[u] >> (v[u] & 31)) & 1u));
            if (keep && lane == 0) s_mask[g + u] = m;
            cnt += __popc(m);
        }
    }
    for (; g < g1; ++g) {
        const uint32_t m = group_mask(g);
        if (keep && lane == 0) s_mask[g] = m;
        cnt += __popc(m);
    }
    if (lane == 0) sw[wid] = cnt;
    __syncthreads();
    // ---- pass 2: exclusive prefix over the warps, then per group from the masks
    const int mine = sw[lane];                             // WS_WARPS == 32
    int run = lane < wid ? mine : 0, total = mine;         // two masked warp sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        run += __shfl_xor_sync(FULL, run, o);
        total += __shfl_xor_sync(FULL, total, o);
    }
    const uint32_t le = 0xffffffffu >> (31 - lane), lt = le >> 1;
    const int L4 = (L + 3) & ~3;                           // the padded tail of the row belongs to this relation too
    for (g = g0; g < g1; ++g) {
        const uint32_t m = keep ? s_mask[g] : group_mask(g);
        const int x = (g << 5) + lane;
        if ((m >> lane) & 1u) pr[run + __popc(m & lt)] = x;
        if (x < L4) nhi[x] = run + __popc(m & le);
        run += __popc(m);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += WS_THREADS) {
        const int rb = x == 0 ? 0 : nhi[x - 1];
        const int target = rb + k;
        F[f_off + x] = target <= total ? lo + pr[target - 1] + 1 : -1;
    }
}

static int window_scan_launch(const int* A, const int* n_acc, const uint32_t* member, int64_t wpr, const int64_t* table,
                              int r_lo, int64_t n_rel_local, int by_order, int* NHI, int* PR, int* F, cudaStream_t s) {
    const size_t smem = size_t(wpr) * 4;
    if (smem <= 84 * 1024) {               // + 24 KB of group masks: two CTAs of 1024 threads per SM still fit
        if (int rc = ensure_dyn_smem((const void*)k_window_scan<true>, smem)) return rc;
        k_window_scan<true><<<(unsigned)n_rel_local, WS_THREADS, smem, s>>>(A, n_acc, member, wpr, table, r_lo, by_order, NHI, PR, F);
    } else {
        k_window_scan<false><<<(unsigned)n_rel_local, WS_THREADS, 0, s>>>(A, n_acc, member, wpr, table, r_lo, by_order, NHI, PR, F);
    }
    return TIPB_OK;
}

// One warp follows o_{r+1} = F_r[o_r - lo_r].  Lane 0 does the dependent lookups out of shared memory: all
// lanes keep staging, with cp.async, the 2 KB of F_{r+AHEAD} around the offset the walk is going to need --
// given o_r, the start of relation r+AHEAD is known to within a few hundred entries (pred[] = expected start
// offsets), far tighter than the bracket itself.  A lookup that falls outside its staged window (never seen
// in practice) simply goes to global memory.
constexpr int WALK_AHEAD = 8;
constexpr int WALK_WIN = 512;  // ints staged per relation
__global__ void __launch_bounds__(32)
k_chain_walk(const int64_t* __restrict__ table, const int* __restrict__ F, int n_rel, int* __restrict__ off,
             int* __restrict__ chain_out, int* __restrict__ status) {
    extern __shared__ int4 s_walk4[];
    int* ring = reinterpret_cast<int*>(s_walk4);            // [AHEAD+1][WALK_WIN]
    int* slot_g0 = ring + (WALK_AHEAD + 1) * WALK_WIN;      // [AHEAD+1] first absolute F index held by a slot
    int* s_tab = slot_g0 + 16;                              // per relation: lo, W, f_off (fits 32 bits), k, pred
    const int lane = threadIdx.x;
    for (int i = lane; i < n_rel; i += 32) {
        const int64_t* tb = table + int64_t(i) * TAB;
        s_tab[5 * i] = int(tb[0]);
        s_tab[5 * i + 1] = int(tb[1]);
        s_tab[5 * i + 2] = int(tb[4]);
        s_tab[5 * i + 3] = int(tb[5]);
        s_tab[5 * i + 4] = int(tb[6]);
    }
    __syncwarp();
    // stage the window of relation `ra`, predicted from offset `o_now` of relation `r_now`; always commits a group
    auto stage = [&](int ra, int r_now, int o_now) {
        if (ra < n_rel) {
            const int xp = o_now + (s_tab[5 * ra + 4] - s_tab[5 * r_now + 4]) - s_tab[5 * ra] - WALK_WIN / 2;
            const int xc = min(max(xp, 0), max(s_tab[5 * ra + 1] - WALK_WIN, 0));
            const int g0 = (s_tab[5 * ra + 2] + xc) & ~3;   // 16-byte aligned absolute index
            const int slot = ra % (WALK_AHEAD + 1);
            if (lane == 0) slot_g0[slot] = g0;
            int* dst = ring + slot * WALK_WIN;
            for (int c = lane; c < WALK_WIN / 4; c += 32) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + 4 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(F + g0 + 4 * c));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    for (int ra = 0; ra < WALK_AHEAD; ++ra) stage(ra, 0, 0);
    int o = 0;
    int r = 0;
    int fail = 0;
    for (; r < n_rel; ++r) {
        stage(r + WALK_AHEAD, r, o);
        asm volatile("cp.async.wait_group %0;" ::"n"(WALK_AHEAD));
        __syncwarp();
        if (lane == 0) {
            off[r] = o;
            if (s_tab[5 * r + 3] != 0) {
                const int x = o - s_tab[5 * r];
                if (x < 0 || x >= s_tab[5 * r + 1]) {
                    fail = NEG_STATUS_BRACKET_MISS;
                } else {
                    const int g = s_tab[5 * r + 2] + x;
                    const int slot = r % (WALK_AHEAD + 1);
                    const int d = g - slot_g0[slot];
                    const int nxt = (d >= 0 && d < WALK_WIN) ? ring[slot * WALK_WIN + d] : F[g];
                    if (nxt < 0) fail = NEG_STATUS_OUT_OF_WORDS; else o = nxt;
                }
            }
        }
        fail = __shfl_sync(FULL, fail, 0);
        if (fail) break;
        o = __shfl_sync(FULL, o, 0);
    }
    asm volatile("cp.async.wait_group 0;");
    if (lane == 0) {
        if (fail) atomicOr(status, fail);
        for (int q = r; q < n_rel; ++q) off[q] = -1;  // relations that could not be placed (r == n_rel: none)
        chain_out[0] = o;
    }
}

// Blocked form of the same walk.  The chain o_{r+1} = F_r[o_r - lo_r] is a composition of 861 table lookups, each
// an L2 round trip when followed one after the other.  Cut the relations into blocks of CHAIN_BLOCK: stage 1
// composes every block for ALL of its first relation's candidate offsets at once (one thread per candidate:
// CHAIN_BLOCK dependent lookups, thousands of candidates in flight), stage 2 walks block to block
// (n_rel / CHAIN_BLOCK lookups) and then lets one thread per block replay its block from the now known start to
// emit off[].  Depth: 2 * CHAIN_BLOCK + n_rel / CHAIN_BLOCK lookups instead of n_rel.
// G[f_off_{r0} + x] = offset of relation r0 + CHAIN_BLOCK if relation r0 starts at lo_{r0} + x; -1: out of words,
// -2: left a bracket.
constexpr int CHAIN_BLOCK = 30;
constexpr int CHAIN_MAX_BLOCKS = 1024;

struct ChainStep { int o; int fail; };
// t = (lo, W, f_off, k) of one relation, staged in shared memory as four ints
__device__ __forceinline__ ChainStep chain_advance(const int4 t, const int* __restrict__ F, int o) {
    ChainStep st{o, 0};
    if (t.w != 0) {
        const int x = o - t.x;
        if (x < 0 || x >= t.y) {
            st.fail = NEG_STATUS_BRACKET_MISS;
        } else {
            const int nxt = F[t.z + x];
            if (nxt < 0) st.fail = NEG_STATUS_OUT_OF_WORDS; else st.o = nxt;
        }
    }
    return st;
}
__device__ __forceinline__ int4 chain_row(const int64_t* __restrict__ table, int r) {
    const int64_t* tb = table + int64_t(r) * TAB;
    return make_int4(int(tb[0]), int(tb[1]), int(tb[4]), int(tb[5]));
}

__global__ void __launch_bounds__(256)
k_chain_blocks(const int64_t* __restrict__ table, const int* __restrict__ F, int r_lo, int n_rel, int* __restrict__ G) {
    // blocks of CHAIN_BLOCK relations counted from r_lo; n_rel = end of the range (a rank's shard, or everything)
    __shared__ int4 s_row[CHAIN_BLOCK];
    const int r0 = r_lo + blockIdx.y * CHAIN_BLOCK;
    const int nr = min(CHAIN_BLOCK, n_rel - r0);
    if (threadIdx.x < nr) s_row[threadIdx.x] = chain_row(table, r0 + threadIdx.x);
    __syncthreads();
    const int lo = s_row[0].x, W = s_row[0].y, f_off = s_row[0].z;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        int o = lo + x;
        int res = 0;
        for (int q = 0; q < nr; ++q) {
            const ChainStep st = chain_advance(s_row[q], F, o);
            if (st.fail) { res = st.fail == NEG_STATUS_OUT_OF_WORDS ? -1 : -2; break; }
            o = st.o;
        }
        G[f_off + x] = res < 0 ? res : o;
    }
}

__global__ void __launch_bounds__(CHAIN_MAX_BLOCKS)
k_chain_stitch(const int64_t* __restrict__ table, const int* __restrict__ F, const int* __restrict__ G, int n_rel,
               int* __restrict__ off, int* __restrict__ chain_out, int* __restrict__ status) {
    extern __shared__ int4 s_tab4[];  // [n_rel]
    __shared__ int s_start[CHAIN_MAX_BLOCKS + 1];
    const int nb = (n_rel + CHAIN_BLOCK - 1) / CHAIN_BLOCK;
    for (int r = threadIdx.x; r < n_rel; r += blockDim.x) s_tab4[r] = chain_row(table, r);
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        int b = 0;
        for (; b < nb; ++b) {
            s_start[b] = o;
            const int4 t0 = s_tab4[b * CHAIN_BLOCK];
            const int x = o - t0.x;
            if (x < 0 || x >= t0.y) {
                // a block whose first relation is empty never consults its bracket: replay it lookup by lookup
                int oo = o, bad = 0;
                for (int r = b * CHAIN_BLOCK; r < min((b + 1) * CHAIN_BLOCK, n_rel); ++r) {
                    const ChainStep st = chain_advance(s_tab4[r], F, oo);
                    if (st.fail) { bad = 1; break; }
                    oo = st.o;
                }
                if (bad) { ++b; break; }
                o = oo;
                continue;
            }
            const int nxt = G[t0.z + x];
            if (nxt < 0) { ++b; break; }
            o = nxt;
        }
        for (int q = b; q < nb; ++q) s_start[q] = -1;  // blocks after a failure
        s_start[nb] = o;
    }
    __syncthreads();
    const int b = threadIdx.x;
    if (b >= nb) return;
    int o = s_start[b];
    const int r0 = b * CHAIN_BLOCK, r1 = min(r0 + CHAIN_BLOCK, n_rel);
    if (o < 0) {
        for (int r = r0; r < r1; ++r) off[r] = -1;
        return;
    }
    for (int r = r0; r < r1; ++r) {
        const ChainStep st = chain_advance(s_tab4[r], F, o);
        if (st.fail) {
            atomicOr(status, st.fail);
            for (int q = r; q < r1; ++q) off[q] = -1;
            chain_out[0] = o;  // the last offset that could be placed
            return;
        }
        off[r] = o;
        o = st.o;
    }
    if (b == nb - 1) chain_out[0] = o;
}

// rounds 0 and 1, one thread per draw.  A CTA takes MAT_TRIPS x 256 consecutive draws: the relation of its first draw
// is found once (the first version searched per 256 draws, ten dependent loads in front of every CTA -- 27 waves of
// CTAs that lived for that search: 94 us for 133 MB of traffic) by a 32-way search, and every thread walks forward
// from it (its draws increase).
constexpr int MAT_TRIPS = 16;

__global__ void __launch_bounds__(256)
k_materialize_main(const int* __restrict__ A, const uint32_t* __restrict__ member, int64_t words_per_rel,
                   const int64_t* __restrict__ range_list, const int64_t* __restrict__ table,
                   const int* __restrict__ off, const int* __restrict__ NHI, int n_rel, int n_nodes, int64_t n_edges,
                   int r_lo, int64_t e_lo, int64_t e_hi, int64_t* __restrict__ out, uint32_t* __restrict__ packed) {
    // draws [e_lo, e_hi) of the relations [r_lo, ...) (a rank's shard; the outputs are indexed from e_lo and hold
    // n_edges = e_hi - e_lo pairs)
    __shared__ int s_first;
    const int64_t e0 = e_lo + int64_t(blockIdx.x) * (MAT_TRIPS * 256);
    if (threadIdx.x < 32) {
        // last relation whose range starts at or before e0: 32-way narrowing of [lo_r, hi_r] (two rounds for 861)
        int lo_r = 0, hi_r = n_rel - 1;
        while (lo_r < hi_r) {
            const int step = (hi_r - lo_r + 31) / 32;
            const int cand = min(lo_r + (int(threadIdx.x) + 1) * step, hi_r);
            const unsigned m = __ballot_sync(FULL, range_list[2 * cand] <= e0);      // monotone over the lanes
            if (m == 0u) {
                hi_r = min(lo_r + step - 1, hi_r);
            } else {
                const int top = 31 - __clz(m);
                lo_r = min(lo_r + (top + 1) * step, hi_r);
                hi_r = min(lo_r + step - 1, hi_r);
            }
        }
        if (threadIdx.x == 0) s_first = lo_r;
    }
    __syncthreads();
    int r = s_first;
    const float fn = float(n_nodes);
#pragma unroll 1
    for (int it = 0; it < MAT_TRIPS; ++it) {
        const int64_t e = e0 + it * 256 + threadIdx.x;
        if (e >= e_hi) return;
        while (r + 1 < n_rel && range_list[2 * (r + 1)] <= e) ++r;
        const int64_t start = range_list[2 * r];
        if (e >= range_list[2 * r + 1]) continue;
        const int o = off[r];
        if (o < 0) continue;
        const int64_t* tb = table + int64_t(r) * TAB;
        const int lo = int(tb[0]), k = int(tb[5]);
        const int* nhi = NHI + tb[3];
        const int i = int(e - start);
        int p = A[o + i];
        // membership of a window entry is what k_window_scan already decided: entry x is a positive pair iff the
        // non-member count does not move at x.  Two coalesced loads instead of a random bitmap probe per draw.
        const int x0 = o - lo, x = x0 + i;
        const int cur = nhi[x], prev = x == 0 ? 0 : nhi[x - 1];
        if (cur == prev) {
            const int rb = x0 == 0 ? 0 : nhi[x0 - 1];
            const int hits_incl = (i + 1) - (cur - rb);
            p = A[o + k + hits_incl - 1];  // the (hits_incl)-th value of round 1
        }
        write_pair(out, packed, n_edges, e - e_lo, p, n_nodes, fn);
    }
}

// rounds >= 2, one CTA per relation
__global__ void __launch_bounds__(256)
k_materialize_fixup(const int* __restrict__ A, const uint32_t* __restrict__ member, int64_t words_per_rel,
                    const int64_t* __restrict__ range_list, const int64_t* __restrict__ table,
                    const int* __restrict__ off, const int* __restrict__ NHI, int n_rel, int n_nodes, int64_t n_edges,
                    int r_lo, int64_t e_lo, int by_order, int64_t* __restrict__ out, uint32_t* __restrict__ packed) {
    // unsharded launches take the relations longest window first (table column 7), as k_window_scan does
    const int r = by_order ? int(table[int64_t(blockIdx.x) * TAB + 7]) : r_lo + blockIdx.x;
    if (r >= n_rel) return;
    const int o = off[r];
    const int64_t* tb = table + int64_t(r) * TAB;
    const int lo = int(tb[0]), k = int(tb[5]);
    if (o < 0 || k == 0) return;
    const int* nhi = NHI + tb[3];
    const int64_t start = range_list[2 * r] - e_lo;
    const int x0 = o - lo;
    const int rb = x0 == 0 ? 0 : nhi[x0 - 1];
    int c_prev = k - (nhi[x0 + k - 1] - rb);  // size of round 1 = hits of round 0
    int s_prev = o + k;                        // round 1 = A[s_prev, s_prev + c_prev)
    const float fn = float(n_nodes);
    // Every round is a run of consecutive stream positions inside the window, so the rank of a hit inside its
    // round follows from the non-member prefix counts; rounds are applied in order (later rounds overwrite).
    while (c_prev > 0) {
        const int s_cur = s_prev + c_prev;
        const int xs = s_prev - lo;              // >= 1
        const int nb = nhi[xs - 1];
        for (int p = threadIdx.x; p < c_prev; p += blockDim.x) {
            const int cur = nhi[xs + p];
            if (cur == nhi[xs + p - 1]) {          // a positive pair (xs >= 1): the non-member count did not move
                const int t = (p + 1) - (cur - nb) - 1;
                write_pair(out, packed, n_edges, start + p, A[s_cur + t], n_nodes, fn);  // perm[rest] = tmp; rest indexes tmp_{q-1}
            }
        }
        const int c = c_prev - (nhi[xs + c_prev - 1] - nb);
        __syncthreads();
        s_prev = s_cur;
        c_prev = c;
    }
}

// ------------------------------------------------------------------------------------------------
// Relation-sharded form (one process per GPU): a rank scans only the windows of ITS relations
// [r_lo, r_hi) and needs its start offset in the shared accepted stream.  The chain is a composition of maps, so
// every rank composes its own relations into ONE table (k_chain_rank_compose: start offset of relation r_lo -> start
// offset of relation r_hi, for every candidate offset of r_lo's bracket), the ranks all-gather these tables (a few KB
// each; the only exchange step of the sampler) and every rank walks the `world` tables (k_chain_resolve).
// Codes in the tables: >= 0 offset, -1 out of words, -2 left a bracket.
__device__ __forceinline__ int chain_walk_range(const int64_t* __restrict__ table, const int* __restrict__ F,
                                                const int* __restrict__ G, int r_begin, int r_end, int block_origin,
                                                int o) {
    // relations [r_begin, r_end); blocks of CHAIN_BLOCK counted from block_origin; G is used at block starts whose
    // bracket holds o, otherwise the block is replayed lookup by lookup.  Returns the offset or a negative code.
    int r = r_begin;
    while (r < r_end) {
        const int4 t = chain_row(table, r);
        const bool at_block_start = ((r - block_origin) % CHAIN_BLOCK) == 0;
        const int blk_end = min(r + CHAIN_BLOCK, r_end);
        const int x = o - t.x;
        if (at_block_start && x >= 0 && x < t.y) {
            const int nxt = G[t.z + x];
            if (nxt < 0) return nxt;
            o = nxt;
            r = blk_end;
        } else {
            const ChainStep st = chain_advance(t, F, o);
            if (st.fail) return st.fail == NEG_STATUS_OUT_OF_WORDS ? -1 : -2;
            o = st.o;
            ++r;
        }
    }
    return o;
}

__global__ void __launch_bounds__(256)
k_chain_rank_compose(const int64_t* __restrict__ table, const int* __restrict__ F, const int* __restrict__ G, int r_lo,
                     int r_hi, int w_max, int* __restrict__ rank_table) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w_max) return;
    if (r_lo >= r_hi) { rank_table[x] = 0; return; }          // a rank without relations: never consulted
    const int4 t0 = chain_row(table, r_lo);
    rank_table[x] = x < t0.y ? chain_walk_range(table, F, G, r_lo, r_hi, r_lo, t0.x + x) : -2;
}

// one CTA: thread 0 walks the ranks' tables, then this rank's blocks; thread b replays block b to emit off[]
__global__ void __launch_bounds__(CHAIN_MAX_BLOCKS)
k_chain_resolve(const int64_t* __restrict__ table, const int* __restrict__ F, const int* __restrict__ G,
                const int* __restrict__ all_tables, const int* __restrict__ first_rel, int world, int rank, int w_max,
                int* __restrict__ off, int* __restrict__ chain_out, int* __restrict__ status) {
    __shared__ int s_start[CHAIN_MAX_BLOCKS + 1];
    __shared__ int s_fail;
    const int r_lo = first_rel[rank], r_hi = first_rel[rank + 1];
    const int nb = (r_hi - r_lo + CHAIN_BLOCK - 1) / CHAIN_BLOCK;
    if (threadIdx.x == 0) {
        int o = 0, mine = 0, fail = 0;
        for (int k = 0; k < world && !fail; ++k) {
            if (k == rank) mine = o;
            const int a = first_rel[k], b = first_rel[k + 1];
            if (a >= b) continue;
            const int4 t0 = chain_row(table, a);
            const int x = o - t0.x;
            if (x < 0 || x >= t0.y || x >= w_max) { fail = NEG_STATUS_BRACKET_MISS; break; }
            const int nxt = all_tables[int64_t(k) * w_max + x];
            if (nxt < 0) { fail = nxt == -1 ? NEG_STATUS_OUT_OF_WORDS : NEG_STATUS_BRACKET_MISS; break; }
            o = nxt;
        }
        s_fail = fail;
        if (fail) {
            atomicOr(status, fail);
        } else {
            chain_out[0] = o;                                   // accepted values consumed by ALL ranks' relations
            int oo = mine;
            for (int b = 0; b < nb; ++b) {                      // start offsets of this rank's blocks
                s_start[b] = oo;
                const int rb = r_lo + b * CHAIN_BLOCK;
                oo = chain_walk_range(table, F, G, rb, min(rb + CHAIN_BLOCK, r_hi), r_lo, oo);
                if (oo < 0) { oo = 0; }                         // cannot happen: the same lookups composed the rank table
            }
        }
    }
    __syncthreads();
    const int b = threadIdx.x;
    if (b >= nb) return;
    const int r0 = r_lo + b * CHAIN_BLOCK, r1 = min(r0 + CHAIN_BLOCK, r_hi);
    if (s_fail) {
        for (int r = r0; r < r1; ++r) off[r] = -1;
        return;
    }
    int o = s_start[b];
    for (int r = r0; r < r1; ++r) {
        off[r] = o;
        const ChainStep st = chain_advance(chain_row(table, r), F, o);
        if (st.fail) {                                           // (unreachable for the same reason)
            atomicOr(status, st.fail);
            for (int q = r + 1; q < r1; ++q) off[q] = -1;
            return;
        }
        o = st.o;
    }
}

// ================================================================================================
// exact (sequential) path: no brackets, used when the fast path reports a bracket miss
// ================================================================================================
// rounds[round_ptr[r] + q] = (start, len) in the accepted stream, q < n_rounds[r]
__global__ void __launch_bounds__(1024)
k_chain_exact(const int* __restrict__ A, const int* __restrict__ n_accepted_ptr, const uint32_t* __restrict__ member,
              int64_t words_per_rel, const int64_t* __restrict__ range_list, int n_rel, int round_cap,
              int* __restrict__ rounds, int* __restrict__ round_ptr, int* __restrict__ n_rounds,
              int* __restrict__ chain_out, int* __restrict__ status) {
    __shared__ int sw[32];
    __shared__ int s_total;
    const int n_acc = *n_accepted_ptr;
    int base = 0;
    int used = 0;
    for (int r = 0; r < n_rel; ++r) {
        const uint32_t* bits = member + int64_t(r) * words_per_rel;
        int n = int(range_list[2 * r + 1] - range_list[2 * r]);
        int q = 0;
        if (threadIdx.x == 0) round_ptr[r] = used;
        while (n > 0) {
            if (base + n > n_acc || used >= round_cap) {
                if (threadIdx.x == 0) atomicOr(status, base + n > n_acc ? NEG_STATUS_OUT_OF_WORDS : NEG_STATUS_TOO_MANY_ROUNDS);
                n = 0;
                base = n_acc;
                break;
            }
            if (threadIdx.x == 0) {
                rounds[2 * used] = base;
                rounds[2 * used + 1] = n;
            }
            ++used;
            int c = 0;
            for (int i = threadIdx.x; i < n; i += 1024) c += is_member(bits, A[base + i]) ? 1 : 0;
            c = warp_sum_i(c);
            if (lane_id() == 0) sw[warp_id()] = c;
            __syncthreads();
            if (warp_id() == 0) {
                int v = sw[lane_id()];
                v = warp_sum_i(v);
                if (lane_id() == 0) s_total = v;
            }
            __syncthreads();
            base += n;
            n = s_total;
            ++q;
            __syncthreads();
        }
        if (threadIdx.x == 0) n_rounds[r] = q;
    }
    if (threadIdx.x == 0) chain_out[0] = base;
}

__global__ void __launch_bounds__(256)
k_materialize_exact(const int* __restrict__ A, const uint32_t* __restrict__ member, int64_t words_per_rel,
                    const int64_t* __restrict__ range_list, const int* __restrict__ rounds,
                    const int* __restrict__ round_ptr, const int* __restrict__ n_rounds, int n_nodes, int64_t n_edges,
                    int* __restrict__ perm, int64_t* __restrict__ out, uint32_t* __restrict__ packed) {
    __shared__ int sw[33];
    __shared__ int s_carry;
    const int r = blockIdx.x;
    const uint32_t* bits = member + int64_t(r) * words_per_rel;
    const int64_t start = range_list[2 * r];
    const int k = int(range_list[2 * r + 1] - start);
    const int nr = n_rounds[r];
    if (k <= 0 || nr <= 0) return;
    int* pr = perm + start;
    const int* rd = rounds + 2 * int64_t(round_ptr[r]);
    const int a0 = rd[0];
    for (int i = threadIdx.x; i < k; i += blockDim.x) pr[i] = A[a0 + i];
    __syncthreads();
    for (int q = 1; q < nr; ++q) {
        const int prev_start = rd[2 * (q - 1)], prev_len = rd[2 * (q - 1) + 1];
        const int cur_start = rd[2 * q];
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < prev_len; base += blockDim.x) {
            const int i = base + threadIdx.x;
            const int hit = (i < prev_len && is_member(bits, A[prev_start + i])) ? 1 : 0;
            const unsigned bal = __ballot_sync(FULL, hit);
            const int in_warp = __popc(bal & ((1u << lane_id()) - 1u));
            if (lane_id() == 0) sw[warp_id()] = __popc(bal);
            __syncthreads();
            if (warp_id() == 0) {
                int v = lane_id() < (blockDim.x >> 5) ? sw[lane_id()] : 0;
                int x = v;
                for (int o = 1; o < 32; o <<= 1) {
                    int y = __shfl_up_sync(FULL, x, o);
                    if (lane_id() >= o) x += y;
                }
                sw[lane_id()] = x - v;
                if (lane_id() == 31) sw[32] = x;
            }
            __syncthreads();
            const int carry = s_carry;
            if (hit) pr[i] = A[cur_start + carry + sw[warp_id()] + in_warp];
            __syncthreads();
            if (threadIdx.x == 0) s_carry = carry + sw[32];
            __syncthreads();
        }
    }
    __syncthreads();
    const float fn = float(n_nodes);
    for (int i = threadIdx.x; i < k; i += blockDim.x) write_pair(out, packed, n_edges, start + i, pr[i], n_nodes, fn);
}

// ================================================================================================
// the call consumed the accepted values [0, consumed): the MT19937 state moves to the word after the last of them.  Its
// position in the raw stream: the chunk whose offset range holds it (binary search over the scanned chunk counts), then
// the chunk's acceptance flags once more (one CTA = one chunk of k_compact).
__global__ void __launch_bounds__(ACC_THREADS)
k_finalize(const uint32_t* __restrict__ U, int64_t n_words, uint32_t mask, uint32_t max_val,
           const int* __restrict__ chunk_off, int n_chunks, const int* __restrict__ chain_out,
           const int* __restrict__ call_status, int* __restrict__ sticky_status, uint32_t* __restrict__ state) {
    __shared__ int64_t s_flat;
    __shared__ int s_chunk, s_target;
    __shared__ int sw[ACC_THREADS / 32];
    if (threadIdx.x == 0) {
        const int failed = *call_status;
        if (failed) atomicOr(sticky_status, failed);   // sticky: only the host clears it
        const int consumed = chain_out[0];
        s_flat = -1;                                   // a failed call consumes nothing
        s_chunk = -1;
        if (!failed && consumed > 0 && consumed <= chunk_off[n_chunks]) {
            int lo = 0, hi = n_chunks - 1;             // last chunk with chunk_off[c] <= consumed - 1
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (chunk_off[mid] <= consumed - 1) lo = mid; else hi = mid - 1;
            }
            s_chunk = lo;
            s_target = consumed - 1 - chunk_off[lo];   // rank of the value inside its chunk
        }
    }
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk < 0) return;
    uint32_t w[ACC_ITEMS];
    const int64_t base = int64_t(chunk) * ACC_CHUNK + int64_t(threadIdx.x) * ACC_ITEMS;
    const unsigned bits = accept_bits(U, base, n_words, int64_t(state[MT_N]), mask, max_val, w);
    const int c = __popc(bits);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, incl, o);
        if (lane_id() >= o) incl += y;
    }
    if (lane_id() == 31) sw[warp_id()] = incl;
    __syncthreads();
    int before = incl - c;
    for (int k = 0; k < warp_id(); ++k) before += sw[k];
    const int t = s_target - before;                   // rank inside this thread's run
    if (t >= 0 && t < c) {
        int seen = 0;
#pragma unroll
        for (int i = 0; i < ACC_ITEMS; ++i) {
            if ((bits >> i) & 1u) {
                if (seen == t) s_flat = base + i + 1;
                ++seen;
            }
        }
    }
    __syncthreads();
    const int64_t flat = s_flat;  // index (in U) of the next unread word
    if (flat < 0) return;
    const int64_t block = (flat - 1) / MT_N;  // numpy keeps pos in [1,624] once a block has been touched
    for (int i = threadIdx.x; i < MT_N; i += blockDim.x) state[i] = U[block * MT_N + i];
    if (threadIdx.x == 0) state[MT_N] = uint32_t(flat - block * MT_N);
}

__global__ void k_bitmap_build(const int64_t* __restrict__ pos_edge_index, const int64_t* __restrict__ range_list,
                               int64_t n_edges, int n_nodes, int n_rel, int64_t words_per_rel, int r_lo, int64_t e_lo,
                               int64_t e_hi, uint32_t* __restrict__ member) {
    // edges [e_lo, e_hi) = the relations [r_lo, ...) whose bitmaps `member` holds (a rank's shard; 0 / n_edges: all)
    int64_t e = e_lo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= e_hi) return;
    int lo = 0, hi = n_rel - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (range_list[2 * mid] <= e) lo = mid; else hi = mid - 1;
    }
    if (!(range_list[2 * lo] <= e && e < range_list[2 * lo + 1])) return;
    const int64_t a = pos_edge_index[e], b = pos_edge_index[n_edges + e];
    if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) return;
    const int64_t key = a * n_nodes + b;
    atomicOr(&member[int64_t(lo - r_lo) * words_per_rel + (key >> 5)], 1u << (key & 31));
}

__global__ void __launch_bounds__(256)
k_bitmap_popcount(const uint32_t* __restrict__ member, int64_t words_per_rel, int* __restrict__ counts) {
    __shared__ int sw[8];
    const uint32_t* bits = member + int64_t(blockIdx.x) * words_per_rel;
    int c = 0;
    for (int64_t i = threadIdx.x; i < words_per_rel; i += 256) c += __popc(bits[i]);
    c = warp_sum_i(c);
    if (lane_id() == 0) sw[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sw[w];
        counts[blockIdx.x] = t;
    }
}

static int64_t bitmap_words(int64_t n_nodes) { return (n_nodes * n_nodes + 31) / 32; }

struct NegWs {
    int *flags, *A, *NHI, *PR, *F, *G, *off, *chain_out, *call_status;
    int *perm, *rounds, *round_ptr, *n_rounds;
    int round_cap;
    void* scan_ws;
};
static size_t neg_ws_layout(int64_t n_edges, int64_t n_rel, int64_t n_words, int64_t sum_l, int64_t sum_w, void* base,
                            NegWs* w) {
    Carver c(base);
    NegWs t;
    t.flags = c.take<int>(n_words + 2);
    t.A = c.take<int>(n_words + 2);
    t.NHI = c.take<int>(sum_l + 1);
    t.PR = c.take<int>(sum_l + 1);
    t.F = c.take<int>(sum_w + 1024);  // slack: the chain walk stages 2 KB windows with 16-byte copies
    t.G = c.take<int>(sum_w + 1024);
    t.off = c.take<int>(n_rel + 1);
    t.chain_out = c.take<int>(4);
    t.call_status = c.take<int>(4);
    t.perm = c.take<int>(n_edges + 1);
    t.round_cap = int(n_rel * 8 + 65536);  // a relation whose pairs cover 99% of the cells needs ~1500 rounds
    t.rounds = c.take<int>(size_t(t.round_cap) * 2);
    t.round_ptr = c.take<int>(n_rel + 1);
    t.n_rounds = c.take<int>(n_rel + 1);
    t.scan_ws = c.take<char>(scan_ws_bytes(n_words + 2));
    if (w) *w = t;
    return c.used() + 256;
}

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_mt19937_seed(uint32_t* mt_state, uint32_t seed, void* stream) {
    TIPB_CHECK_ARG(mt_state, "mt19937_seed: NULL state");
    k_mt_seed<<<1, 32, 0, (cudaStream_t)stream>>>(mt_state, seed);
    TIPB_CHECK_LAUNCH("mt19937_seed");
    return TIPB_OK;
}

int64_t tipb_mt19937_stream_words(int64_t n_new) {
    const int64_t q = 2 * MT_LAG;
    return (n_new + q - 1) / q * q;
}

int tipb_mt19937_generate(const uint32_t* mt_state, uint32_t* stream_words, int64_t n_new, void* stream) {
    TIPB_CHECK_ARG(mt_state && stream_words, "mt19937_generate: NULL argument");
    TIPB_CHECK_ARG(n_new > 0 && n_new % (2 * MT_LAG) == 0 && n_new < (int64_t(1) << 31) - 4096,
                   "mt19937_generate: n_new must be a positive multiple of 454 (use tipb_mt19937_stream_words)");
    k_mt_generate<<<1, 256, 0, (cudaStream_t)stream>>>(mt_state, stream_words, (int)n_new);
    TIPB_CHECK_LAUNCH("mt19937_generate");
    return TIPB_OK;
}

size_t tipb_neg_bitmap_bytes(int64_t n_nodes, int64_t n_rel) { return size_t(bitmap_words(n_nodes)) * n_rel * 4; }

int tipb_neg_bitmap_build(const int64_t* pos_edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_nodes,
                          int64_t n_rel, uint32_t* member, int32_t* popcount, void* stream) {
    return tipb_neg_bitmap_build_range(pos_edge_index, range_list, n_edges, n_nodes, n_rel, 0, n_rel, 0, n_edges, member,
                                       popcount, stream);
}

int tipb_neg_bitmap_build_range(const int64_t* pos_edge_index, const int64_t* range_list, int64_t n_edges,
                                int64_t n_nodes, int64_t n_rel, int64_t r_lo, int64_t r_hi, int64_t e_lo, int64_t e_hi,
                                uint32_t* member, int32_t* popcount, void* stream) {
    TIPB_CHECK_ARG(range_list && member && popcount && (n_edges == 0 || pos_edge_index), "neg_bitmap_build: NULL argument");
    TIPB_CHECK_ARG(n_nodes > 0 && n_nodes <= 46340, "neg_bitmap_build: n_nodes^2 must fit in int32");
    TIPB_CHECK_ARG(0 <= r_lo && r_lo <= r_hi && r_hi <= n_rel && 0 <= e_lo && e_lo <= e_hi && e_hi <= n_edges,
                   "neg_bitmap_build: bad relation / edge range");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_local = r_hi - r_lo;
    TIPB_CHECK_CUDA(cudaMemsetAsync(member, 0, tipb_neg_bitmap_bytes(n_nodes, n_local > 0 ? n_local : 1), s));
    if (e_hi > e_lo)
        k_bitmap_build<<<(unsigned)ceil_div(e_hi - e_lo, 256), 256, 0, s>>>(pos_edge_index, range_list, n_edges, (int)n_nodes,
                                                                            (int)n_rel, bitmap_words(n_nodes), (int)r_lo,
                                                                            e_lo, e_hi, member);
    if (n_local > 0) k_bitmap_popcount<<<(unsigned)n_local, 256, 0, s>>>(member, bitmap_words(n_nodes), popcount);
    TIPB_CHECK_LAUNCH("neg_bitmap_build");
    return TIPB_OK;
}

// Host-only (no CUDA call): brackets of the per-relation start offsets from a negative-binomial model.
//   relation r needs k_r non-members; with member density d_r the number of extra draws has mean k d/(1-d)
//   and variance k d/(1-d)^2.  Offsets accumulate over relations, so do means and variances.
int tipb_neg_table_build(const int64_t* range_list_host, const int32_t* popcount_host, int64_t n_rel, int64_t n_nodes,
                         double z_sigma, int64_t* table_host, int64_t* totals_host) {
    TIPB_CHECK_ARG(range_list_host && popcount_host && table_host && totals_host, "neg_table_build: NULL argument");
    const double cells = double(n_nodes) * double(n_nodes);
    const int64_t slack = 64;
    double mean = 0.0, var = 0.0;
    int64_t ksum = 0, sum_l = 0, sum_w = 0, max_index = 0;
    for (int64_t r = 0; r < n_rel; ++r) {
        const int64_t k = range_list_host[2 * r + 1] - range_list_host[2 * r];
        TIPB_CHECK_ARG(k >= 0, "neg_table_build: negative relation size");
        double d = double(popcount_host[r]) / cells;
        if (d > 0.98) d = 0.98;
        const double dev = z_sigma * sqrt(var);
        int64_t lo_x = int64_t(floor(mean - dev)) - slack;
        if (lo_x < 0) lo_x = 0;
        const int64_t hi_x = int64_t(ceil(mean + dev)) + slack;
        const int64_t lo = ksum + lo_x, W = hi_x - lo_x + 1;
        const double ek = double(k) * d / (1.0 - d), vk = double(k) * d / ((1.0 - d) * (1.0 - d));
        const int64_t L = k == 0 ? 0 : W + k + int64_t(ceil(ek + z_sigma * sqrt(vk))) + slack;
        int64_t* tb = table_host + r * TAB;
        tb[0] = lo; tb[1] = W; tb[2] = L; tb[3] = sum_l; tb[4] = sum_w; tb[5] = k;
        tb[6] = ksum + int64_t(llround(mean));
        sum_l += (L + 3) & ~int64_t(3);       // rows of NHI / PR start 16-byte aligned (vector stores in k_window_scan)
        sum_w += W;
        if (lo + L > max_index) max_index = lo + L;
        if (lo + W > max_index) max_index = lo + W;
        ksum += k;
        mean += ek;
        var += vk;
    }
    TIPB_CHECK_ARG(max_index < (int64_t(1) << 31) - 4096 && sum_l < (int64_t(1) << 31) && sum_w < (int64_t(1) << 31),
                   "neg_table_build: edge set too large for 32-bit stream offsets");
    {   // column 7: relations by window length, longest first (stable)
        std::vector<int64_t> order(static_cast<size_t>(n_rel));
        for (int64_t r = 0; r < n_rel; ++r) order[size_t(r)] = r;
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return table_host[a * TAB + 2] > table_host[b * TAB + 2]; });
        for (int64_t b = 0; b < n_rel; ++b) table_host[b * TAB + 7] = order[size_t(b)];
    }
    totals_host[0] = sum_l;
    totals_host[1] = sum_w;
    totals_host[2] = max_index;                        // accepted values the windows may touch
    totals_host[3] = ksum + int64_t(ceil(mean));       // expected number of accepted values consumed
    return TIPB_OK;
}

size_t tipb_neg_sample_workspace_bytes(int64_t n_edges, int64_t n_rel, int64_t n_words, int64_t sum_l, int64_t sum_w) {
    return neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, nullptr, nullptr);
}

int tipb_neg_sample(uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words, const uint32_t* member,
                    const int64_t* range_list, const int64_t* table, int64_t sum_l, int64_t sum_w, int64_t n_edges,
                    int64_t n_nodes, int64_t n_rel, int exact_mode, int64_t* neg_edge_index, uint32_t* neg_packed,
                    int32_t* status, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(mt_state && stream_words && member && range_list && (neg_edge_index || neg_packed) && status && ws,
                   "neg_sample: NULL argument");
    TIPB_CHECK_ARG(!neg_packed || n_nodes <= 65535, "neg_sample: the packed output holds 16-bit node ids");
    TIPB_CHECK_ARG(exact_mode || table, "neg_sample: the fast path needs the bracket table");
    TIPB_CHECK_ARG(n_nodes > 1 && n_nodes <= 46340, "neg_sample: n_nodes must be in [2, 46340]");
    TIPB_CHECK_ARG(n_words > MT_N && n_words < (int64_t(1) << 31) - 4096, "neg_sample: bad stream length");
    TIPB_CHECK_ARG(n_rel > 0 && n_rel * 20 <= 180 * 1024, "neg_sample: n_rel out of range");
    TIPB_CHECK_ARG((reinterpret_cast<uintptr_t>(stream_words) & 15) == 0, "neg_sample: stream_words must be 16-byte aligned");
    TIPB_CHECK_ARG(ws_bytes >= neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, nullptr, nullptr),
                   "neg_sample: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    NegWs w;
    neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, ws, &w);
    const uint32_t max_val = uint32_t(n_nodes * n_nodes - 1);
    uint32_t mask = 1;
    while (mask < max_val) mask = (mask << 1) | 1u;
    const int T = 256;
    const int64_t wpr = bitmap_words(n_nodes);
    int rc;

    // `status` is STICKY: failure bits of every call are OR-ed into it and only the caller clears it (after reading
    // it on the host).  The bits of THIS call are collected in a workspace word; k_finalize folds them into `status`
    // and leaves the MT19937 state untouched when the call failed (the caller reruns it from the same state).
    int* call_status = w.call_status;
    TIPB_CHECK_CUDA(cudaMemsetAsync(call_status, 0, sizeof(int32_t), s));
    const int64_t n_chunks = ceil_div(n_words, ACC_CHUNK);
    k_accept_count<<<(unsigned)n_chunks, ACC_THREADS, 0, s>>>(stream_words, mt_state + MT_N, n_words, mask, max_val, w.flags);
    if ((rc = exclusive_scan_i32(w.flags, w.flags, n_chunks, w.scan_ws, s))) return rc;
    k_compact<<<(unsigned)n_chunks, ACC_THREADS, 0, s>>>(stream_words, mt_state + MT_N, n_words, mask, max_val, w.flags,
                                                         w.A);
    const int* n_acc = w.flags + n_chunks;
    if (exact_mode) {
        k_chain_exact<<<1, 1024, 0, s>>>(w.A, n_acc, member, wpr, range_list, (int)n_rel, w.round_cap, w.rounds,
                                         w.round_ptr, w.n_rounds, w.chain_out, call_status);
        k_materialize_exact<<<(unsigned)n_rel, 256, 0, s>>>(w.A, member, wpr, range_list, w.rounds, w.round_ptr,
                                                            w.n_rounds, (int)n_nodes, n_edges, w.perm, neg_edge_index, neg_packed);
    } else {
        if ((rc = window_scan_launch(w.A, n_acc, member, wpr, table, 0, n_rel, 1, w.NHI, w.PR, w.F, s))) return rc;
        const int n_blocks = int(ceil_div(n_rel, CHAIN_BLOCK));
        if (n_blocks <= CHAIN_MAX_BLOCKS && size_t(n_rel) * sizeof(int4) <= 160 * 1024) {
            const size_t tsm = size_t(n_rel) * sizeof(int4);
            if ((rc = ensure_dyn_smem((const void*)k_chain_stitch, tsm))) return rc;
            k_chain_blocks<<<dim3(16, (unsigned)n_blocks), T, 0, s>>>(table, w.F, 0, (int)n_rel, w.G);
            k_chain_stitch<<<1, CHAIN_MAX_BLOCKS, tsm, s>>>(table, w.F, w.G, (int)n_rel, w.off, w.chain_out, call_status);
        } else {
            const size_t smem = size_t(n_rel) * 20 + size_t((WALK_AHEAD + 1) * WALK_WIN + 16) * sizeof(int);
            if ((rc = ensure_dyn_smem((const void*)k_chain_walk, smem))) return rc;
            k_chain_walk<<<1, 32, smem, s>>>(table, w.F, (int)n_rel, w.off, w.chain_out, call_status);
        }
        if (n_edges > 0)
            k_materialize_main<<<(unsigned)ceil_div(n_edges, MAT_TRIPS * 256), 256, 0, s>>>(w.A, member, wpr, range_list, table, w.off,
                                                                           w.NHI, (int)n_rel, (int)n_nodes, n_edges, 0,
                                                                           0, n_edges, neg_edge_index, neg_packed);
        k_materialize_fixup<<<(unsigned)n_rel, T, 0, s>>>(w.A, member, wpr, range_list, table, w.off, w.NHI, (int)n_rel,
                                                          (int)n_nodes, n_edges, 0, 0, 1, neg_edge_index, neg_packed);
    }
    k_finalize<<<1, ACC_THREADS, 0, s>>>(stream_words, n_words, mask, max_val, w.flags, (int)n_chunks, w.chain_out, call_status,
                                         status, mt_state);
    TIPB_CHECK_LAUNCH("neg_sample");
    return TIPB_OK;
}

// ---- relation-sharded sampler: begin (own windows + own rank table) ... all-gather by the caller ... end
static int shard_check(const char* who, int64_t n_rel, int64_t r_lo, int64_t r_hi, int64_t w_max) {
    if (!(0 <= r_lo && r_lo <= r_hi && r_hi <= n_rel) || w_max < 1 || w_max > (int64_t(1) << 28)) {
        set_last_error("%s: bad relation range [%lld, %lld) of %lld or w_max", who, (long long)r_lo, (long long)r_hi,
                       (long long)n_rel);
        return TIPB_ERR_INVALID_ARGUMENT;
    }
    if ((r_hi - r_lo + CHAIN_BLOCK - 1) / CHAIN_BLOCK > CHAIN_MAX_BLOCKS) {
        set_last_error("%s: more than %d relations per rank", who, CHAIN_BLOCK * CHAIN_MAX_BLOCKS);
        return TIPB_ERR_UNSUPPORTED;
    }
    return TIPB_OK;
}

int tipb_neg_sample_shard_begin(const uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words,
                                const uint32_t* member_local, const int64_t* table, int64_t sum_l, int64_t sum_w,
                                int64_t n_edges, int64_t n_nodes, int64_t n_rel, int64_t r_lo, int64_t r_hi,
                                int32_t* rank_table, int64_t w_max, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(mt_state && stream_words && member_local && table && rank_table && ws, "neg_sample_shard_begin: NULL argument");
    TIPB_CHECK_ARG(n_nodes > 1 && n_nodes <= 46340, "neg_sample_shard_begin: n_nodes must be in [2, 46340]");
    TIPB_CHECK_ARG(n_words > MT_N && n_words < (int64_t(1) << 31) - 4096, "neg_sample_shard_begin: bad stream length");
    TIPB_CHECK_ARG((reinterpret_cast<uintptr_t>(stream_words) & 15) == 0, "neg_sample_shard_begin: stream_words must be 16-byte aligned");
    TIPB_CHECK_ARG(ws_bytes >= neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, nullptr, nullptr),
                   "neg_sample_shard_begin: workspace too small");
    if (int rc = shard_check("neg_sample_shard_begin", n_rel, r_lo, r_hi, w_max)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    NegWs w;
    neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, ws, &w);
    const uint32_t max_val = uint32_t(n_nodes * n_nodes - 1);
    uint32_t mask = 1;
    while (mask < max_val) mask = (mask << 1) | 1u;
    const int64_t wpr = bitmap_words(n_nodes);
    int rc;
    TIPB_CHECK_CUDA(cudaMemsetAsync(w.call_status, 0, sizeof(int32_t), s));
    const int64_t n_chunks = ceil_div(n_words, ACC_CHUNK);
    k_accept_count<<<(unsigned)n_chunks, ACC_THREADS, 0, s>>>(stream_words, mt_state + MT_N, n_words, mask, max_val, w.flags);
    if ((rc = exclusive_scan_i32(w.flags, w.flags, n_chunks, w.scan_ws, s))) return rc;
    k_compact<<<(unsigned)n_chunks, ACC_THREADS, 0, s>>>(stream_words, mt_state + MT_N, n_words, mask, max_val, w.flags,
                                                         w.A);
    const int* n_acc = w.flags + n_chunks;
    const int64_t n_local = r_hi - r_lo;
    if (n_local > 0) {
        if ((rc = window_scan_launch(w.A, n_acc, member_local, wpr, table, (int)r_lo, n_local, 0, w.NHI, w.PR, w.F, s))) return rc;
        const int n_blocks = int(ceil_div(n_local, CHAIN_BLOCK));
        k_chain_blocks<<<dim3(16, (unsigned)n_blocks), 256, 0, s>>>(table, w.F, (int)r_lo, (int)r_hi, w.G);
    }
    k_chain_rank_compose<<<(unsigned)ceil_div(w_max, 256), 256, 0, s>>>(table, w.F, w.G, (int)r_lo, (int)r_hi, (int)w_max,
                                                                       rank_table);
    TIPB_CHECK_LAUNCH("neg_sample_shard_begin");
    return TIPB_OK;
}

int tipb_neg_sample_shard_end(uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words,
                              const uint32_t* member_local, const int64_t* range_list, const int64_t* table,
                              int64_t sum_l, int64_t sum_w, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                              const int32_t* all_tables, const int32_t* first_rel, int world, int rank, int64_t w_max,
                              int64_t r_lo, int64_t r_hi, int64_t e_lo, int64_t e_hi, int64_t* neg_local,
                              uint32_t* packed_local, int32_t* status, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(mt_state && stream_words && member_local && range_list && table && all_tables && first_rel && status && ws,
                   "neg_sample_shard_end: NULL argument");
    TIPB_CHECK_ARG(neg_local || packed_local, "neg_sample_shard_end: no output");
    TIPB_CHECK_ARG(!packed_local || n_nodes <= 65535, "neg_sample_shard_end: the packed output holds 16-bit node ids");
    TIPB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && 0 <= e_lo && e_lo <= e_hi && e_hi <= n_edges,
                   "neg_sample_shard_end: bad rank / edge range");
    if (int rc = shard_check("neg_sample_shard_end", n_rel, r_lo, r_hi, w_max)) return rc;
    TIPB_CHECK_ARG(ws_bytes >= neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, nullptr, nullptr),
                   "neg_sample_shard_end: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    NegWs w;
    neg_ws_layout(n_edges, n_rel, n_words, sum_l, sum_w, ws, &w);
    const int64_t wpr = bitmap_words(n_nodes);
    const uint32_t max_val = uint32_t(n_nodes * n_nodes - 1);
    uint32_t mask = 1;
    while (mask < max_val) mask = (mask << 1) | 1u;
    // the relation range is read on the device from first_rel; the host passes the matching edge range
    k_chain_resolve<<<1, CHAIN_MAX_BLOCKS, 0, s>>>(table, w.F, w.G, all_tables, first_rel, world, rank, (int)w_max, w.off,
                                                   w.chain_out, w.call_status);
    const int64_t e_local = e_hi - e_lo;      // = the edges of the relations [r_lo, r_hi) = first_rel[rank .. rank + 1]
    if (e_local > 0)
        k_materialize_main<<<(unsigned)ceil_div(e_local, MAT_TRIPS * 256), 256, 0, s>>>(w.A, member_local, wpr, range_list, table, w.off,
                                                                           w.NHI, (int)n_rel, (int)n_nodes, e_local,
                                                                           (int)r_lo, e_lo, e_hi, neg_local, packed_local);
    if (r_hi > r_lo)
        k_materialize_fixup<<<(unsigned)(r_hi - r_lo), 256, 0, s>>>(w.A, member_local, wpr, range_list, table, w.off, w.NHI,
                                                                  (int)r_hi, (int)n_nodes, e_local, (int)r_lo, e_lo, 0,
                                                                  neg_local, packed_local);
    k_finalize<<<1, ACC_THREADS, 0, s>>>(stream_words, n_words, mask, max_val, w.flags, (int)ceil_div(n_words, ACC_CHUNK), w.chain_out,
                                         w.call_status, status, mt_state);
    TIPB_CHECK_LAUNCH("neg_sample_shard_end");
    return TIPB_OK;
}
}
