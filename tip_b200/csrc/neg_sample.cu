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
// Pipeline (all on `stream`, no host synchronisation):
//   k_mt_generate   one CTA runs the MT19937 recurrence S[n+624] = S[n+397] ^ g(S[n], S[n+1]); thread t
//                   owns outputs 624 + t + 227p, so S[n+397] is its own previous value (register) and the
//                   other two operands were written >= 2 phases ago (one barrier per two phases)
//   k_accept_flags / scan / k_compact   masked-rejection as a stream compaction -> accepted stream A
//   k_chain         one CTA walks the relations in order: per retry round, count members of A[...] in the
//                   relation's positive-pair bitmap; records every round's (start, length)
//   k_materialize   one CTA per relation replays the rounds (stable compaction of hit positions) and
//                   writes the int64 [2,E] result
//   k_finalize      advances the caller's MT19937 state to exactly where numpy's would be
#include "common.cuh"

namespace tipb {

constexpr int MT_N = 624, MT_M = 397, MT_LAG = MT_N - MT_M;  // 227
constexpr uint32_t MT_UPPER = 0x80000000u, MT_LOWER = 0x7fffffffu, MT_MATRIX_A = 0x9908b0dfu;
constexpr int RING = 2048;

enum { NEG_STATUS_OUT_OF_WORDS = 1, NEG_STATUS_TOO_MANY_ROUNDS = 2 };

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    uint32_t y = (a & MT_UPPER) | (b & MT_LOWER);
    return (y >> 1) ^ ((b & 1u) ? MT_MATRIX_A : 0u);
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

// U[0..624) = current key; U[624 + i] = i-th further word of the untempered stream, for i < n_new.
__global__ void __launch_bounds__(256) k_mt_generate(const uint32_t* __restrict__ state, uint32_t* __restrict__ U,
                                                     int64_t n_new) {
    __shared__ uint32_t ring[RING];
    const int t = threadIdx.x;
    for (int i = t; i < MT_N; i += blockDim.x) {
        uint32_t v = state[i];
        ring[i] = v;
        U[i] = v;
    }
    __syncthreads();
    const int64_t n_phases = (n_new + MT_LAG - 1) / MT_LAG;
    uint32_t prev = t < MT_LAG ? ring[MT_M + t] : 0u;  // S[397 + t]
    for (int64_t p = 0; p < n_phases; ++p) {
        if (t < MT_LAG) {
            const int64_t n = t + MT_LAG * p;  // produces S[n + 624]
            const uint32_t a = ring[n & (RING - 1)], b = ring[(n + 1) & (RING - 1)];
            const uint32_t v = prev ^ mt_mix(a, b);
            ring[(n + MT_N) & (RING - 1)] = v;
            if (n < n_new) U[n + MT_N] = v;
            prev = v;
        }
        // operands of phase p+1 were produced in phases <= p-1 (S[n], S[n+1] with n+1 <= 227(p+2)-1+1 < 624+227p
        // only when p >= ... ) -- the first phases read the seed block, later ones need data two phases old,
        // except S[n+1] of thread 226 at phase p, which is S[227(p+1)]: produced by thread 0 in phase p+1-3+... ;
        // a barrier after every phase whose successor could read fresh data keeps this simple and safe:
        if ((p & 1) || p < 4) __syncthreads();
    }
}

// word i of the window is a candidate iff it lies at or after the state's read position
__global__ void k_accept_flags(const uint32_t* __restrict__ U, const uint32_t* __restrict__ pos_ptr, int64_t n_words,
                               uint32_t mask, uint32_t max_val, int* __restrict__ flags) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    uint32_t w = mt_temper(U[i]) & mask;
    flags[i] = (i >= int64_t(*pos_ptr) && w <= max_val) ? 1 : 0;
}

__global__ void k_compact(const uint32_t* __restrict__ U, const uint32_t* __restrict__ pos_ptr, int64_t n_words,
                          uint32_t mask, uint32_t max_val, const int* __restrict__ slot, int* __restrict__ A,
                          int* __restrict__ Apos) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    uint32_t w = mt_temper(U[i]) & mask;
    if (i >= int64_t(*pos_ptr) && w <= max_val) {
        int a = slot[i];
        A[a] = int(w);
        Apos[a] = int(i);
    }
}

__device__ __forceinline__ bool is_member(const uint32_t* __restrict__ bits, int key) {
    return (bits[key >> 5] >> (key & 31)) & 1u;
}

// rounds[round_ptr[r] + q] = (start, len) in the accepted stream, q < n_rounds[r] (one flat table for all
// relations, `round_cap` entries); chain_out[0] = accepted values consumed
__global__ void __launch_bounds__(1024)
k_chain(const int* __restrict__ A, const int* __restrict__ n_accepted_ptr, const uint32_t* __restrict__ member,
        int64_t words_per_rel, const int64_t* __restrict__ range_list, int n_rel, int round_cap,
        int* __restrict__ rounds, int* __restrict__ round_ptr, int* __restrict__ n_rounds, int* __restrict__ chain_out,
        int* __restrict__ status) {
    __shared__ int sw[32];
    __shared__ int s_total;
    const int n_acc = *n_accepted_ptr;
    int base = 0;
    int used = 0;  // entries of the flat round table handed out so far
    for (int r = 0; r < n_rel; ++r) {
        const uint32_t* bits = member + int64_t(r) * words_per_rel;
        int n = int(range_list[2 * r + 1] - range_list[2 * r]);
        int q = 0;
        if (threadIdx.x == 0) round_ptr[r] = used;
        while (n > 0) {
            if (base + n > n_acc || used >= round_cap) {
                if (threadIdx.x == 0) atomicOr(status, base + n > n_acc ? NEG_STATUS_OUT_OF_WORDS : NEG_STATUS_TOO_MANY_ROUNDS);
                n = 0;
                base = n_acc;  // poison: every later relation fails the same way
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

// one CTA per relation: replay the rounds, write int64 pairs
__global__ void __launch_bounds__(256)
k_materialize(const int* __restrict__ A, const uint32_t* __restrict__ member, int64_t words_per_rel,
              const int64_t* __restrict__ range_list, const int* __restrict__ rounds, const int* __restrict__ round_ptr,
              const int* __restrict__ n_rounds, int n_nodes, int64_t n_edges, int* __restrict__ perm,
              int64_t* __restrict__ out) {
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
    // round q >= 1: tmp_q = A[rd[2q] ...], positions = ascending hit positions inside tmp_{q-1}
    for (int q = 1; q < nr; ++q) {
        const int prev_start = rd[2 * (q - 1)], prev_len = rd[2 * (q - 1) + 1];
        const int cur_start = rd[2 * q];
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < prev_len; base += blockDim.x) {
            const int i = base + threadIdx.x;
            const int hit = (i < prev_len && is_member(bits, A[prev_start + i])) ? 1 : 0;
            // block-wide exclusive scan of `hit` (ballot inside the warp, then across warps)
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
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const int p = pr[i];
        const float row = __fdiv_rn(__int2float_rn(p), fn);  // float32 true division, as torch does
        out[start + i] = (long long)row;                       // .long(): truncation
        out[n_edges + start + i] = (long long)(p % n_nodes);
    }
}

__global__ void __launch_bounds__(256)
k_finalize(const uint32_t* __restrict__ U, const int* __restrict__ Apos, const int* __restrict__ chain_out,
           uint32_t* __restrict__ state) {
    __shared__ int64_t s_flat;
    if (threadIdx.x == 0) {
        const int consumed = chain_out[0];
        s_flat = consumed > 0 ? int64_t(Apos[consumed - 1]) + 1 : -1;
    }
    __syncthreads();
    const int64_t flat = s_flat;  // index (in U) of the next unread word
    if (flat < 0) return;
    const int64_t block = (flat - 1) / MT_N;  // numpy keeps pos in [1,624] once a block has been touched
    for (int i = threadIdx.x; i < MT_N; i += blockDim.x) state[i] = U[block * MT_N + i];
    if (threadIdx.x == 0) state[MT_N] = uint32_t(flat - block * MT_N);
}

__global__ void k_bitmap_build(const int64_t* __restrict__ pos_edge_index, const int64_t* __restrict__ range_list,
                               int64_t n_edges, int n_nodes, int n_rel, int64_t words_per_rel,
                               uint32_t* __restrict__ member) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    int lo = 0, hi = n_rel - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (range_list[2 * mid] <= e) lo = mid; else hi = mid - 1;
    }
    if (!(range_list[2 * lo] <= e && e < range_list[2 * lo + 1])) return;
    const int64_t a = pos_edge_index[e], b = pos_edge_index[n_edges + e];
    if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) return;
    const int64_t key = a * n_nodes + b;
    atomicOr(&member[int64_t(lo) * words_per_rel + (key >> 5)], 1u << (key & 31));
}

static int64_t bitmap_words(int64_t n_nodes) { return (n_nodes * n_nodes + 31) / 32; }

struct NegWs {
    uint32_t* U;
    int *flags, *A, *Apos, *perm, *rounds, *round_ptr, *n_rounds, *chain_out;
    int round_cap;
    void* scan_ws;
};
static size_t neg_ws_layout(int64_t n_edges, int64_t n_rel, int64_t budget, void* base, NegWs* w) {
    Carver c(base);
    NegWs t;
    t.U = c.take<uint32_t>(budget + 3 * MT_N);
    t.flags = c.take<int>(budget + MT_N + 2);
    t.A = c.take<int>(budget + MT_N + 2);
    t.Apos = c.take<int>(budget + MT_N + 2);
    t.perm = c.take<int>(n_edges + 1);
    t.round_cap = int(n_rel * 8 + 65536);  // a relation whose pairs cover 99% of the cells needs ~1500 rounds
    t.rounds = c.take<int>(size_t(t.round_cap) * 2);
    t.round_ptr = c.take<int>(n_rel);
    t.n_rounds = c.take<int>(n_rel);
    t.chain_out = c.take<int>(4);
    t.scan_ws = c.take<char>(scan_ws_bytes(budget + MT_N + 2));
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

size_t tipb_neg_bitmap_bytes(int64_t n_nodes, int64_t n_rel) { return size_t(bitmap_words(n_nodes)) * n_rel * 4; }

int tipb_neg_bitmap_build(const int64_t* pos_edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_nodes,
                          int64_t n_rel, uint32_t* member, void* stream) {
    TIPB_CHECK_ARG(range_list && member && (n_edges == 0 || pos_edge_index), "neg_bitmap_build: NULL argument");
    TIPB_CHECK_ARG(n_nodes > 0 && n_nodes <= 46340, "neg_bitmap_build: n_nodes^2 must fit in int32");
    cudaStream_t s = (cudaStream_t)stream;
    TIPB_CHECK_CUDA(cudaMemsetAsync(member, 0, tipb_neg_bitmap_bytes(n_nodes, n_rel), s));
    if (n_edges > 0)
        k_bitmap_build<<<(unsigned)ceil_div(n_edges, 256), 256, 0, s>>>(pos_edge_index, range_list, n_edges, (int)n_nodes,
                                                                        (int)n_rel, bitmap_words(n_nodes), member);
    TIPB_CHECK_LAUNCH("neg_bitmap_build");
    return TIPB_OK;
}

size_t tipb_neg_sample_workspace_bytes(int64_t n_edges, int64_t n_rel, int64_t budget_words) {
    return neg_ws_layout(n_edges, n_rel, budget_words, nullptr, nullptr);
}

int tipb_neg_sample(uint32_t* mt_state, const uint32_t* member, const int64_t* range_list, int64_t n_edges,
                    int64_t n_nodes, int64_t n_rel, int64_t budget_words, int64_t* neg_edge_index, int32_t* status,
                    void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(mt_state && member && range_list && neg_edge_index && status && ws, "neg_sample: NULL argument");
    TIPB_CHECK_ARG(n_nodes > 1 && n_nodes <= 46340, "neg_sample: n_nodes must be in [2, 46340]");
    TIPB_CHECK_ARG(budget_words > 0 && budget_words < (int64_t(1) << 31) - 4096, "neg_sample: bad word budget");
    TIPB_CHECK_ARG(ws_bytes >= neg_ws_layout(n_edges, n_rel, budget_words, nullptr, nullptr), "neg_sample: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    NegWs w;
    neg_ws_layout(n_edges, n_rel, budget_words, ws, &w);
    const uint32_t max_val = uint32_t(n_nodes * n_nodes - 1);
    uint32_t mask = 1;
    while (mask < max_val) mask = (mask << 1) | 1u;
    const int64_t n_words = budget_words + MT_N;  // candidate window = current key block + budget new words
    const int T = 256;

    TIPB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    k_mt_generate<<<1, 256, 0, s>>>(mt_state, w.U, budget_words);
    k_accept_flags<<<(unsigned)ceil_div(n_words, T), T, 0, s>>>(w.U, mt_state + MT_N, n_words, mask, max_val, w.flags);
    int rc = exclusive_scan_i32(w.flags, w.flags, n_words, w.scan_ws, s);
    if (rc) return rc;
    k_compact<<<(unsigned)ceil_div(n_words, T), T, 0, s>>>(w.U, mt_state + MT_N, n_words, mask, max_val, w.flags, w.A, w.Apos);
    k_chain<<<1, 1024, 0, s>>>(w.A, w.flags + n_words, member, bitmap_words(n_nodes), range_list, (int)n_rel, w.round_cap,
                              w.rounds, w.round_ptr, w.n_rounds, w.chain_out, status);
    k_materialize<<<(unsigned)n_rel, 256, 0, s>>>(w.A, member, bitmap_words(n_nodes), range_list, w.rounds, w.round_ptr, w.n_rounds,
                                                  (int)n_nodes, n_edges, w.perm, neg_edge_index);
    k_finalize<<<1, 256, 0, s>>>(w.U, w.Apos, w.chain_out, mt_state);
    TIPB_CHECK_LAUNCH("neg_sample");
    return TIPB_OK;
}
}
