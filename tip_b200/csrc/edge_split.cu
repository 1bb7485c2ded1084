// GPU data preparation: the train / test edge split of the reference.
//
//   process_edges(raw_edge_list, p)   src/utils.py:35-65   per relation: np.random.binomial(1, p, n) keep mask,
//                                     kept pairs + their mirror images (to_bidirection, :17-23), cumulative ranges (:26-32)
//   process_prot_edge(pp_net)         data/utils.py:212-229  the same split of the one P-P relation (p = 0.9)
//
// Bit-exact with numpy's legacy global stream: binomial(1, p) is the inversion sampler on min(p, 1-p), which draws ONE
// 53-bit double per edge -- two successive tempered MT19937 words a, b: U = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 -- and
// returns X = [U > qn] with qn = exp(log(1 - min(p, 1-p))) (formed on the host in double, as numpy's C code does);
// keep = 1 - X for p > 0.5, X otherwise.  All draws of all relations are one contiguous run of the stream, so the mask
// is an embarrassingly parallel map over the pre-generated word stream (csrc/mt_jump.cu), the placement is one
// exclusive scan, and the MT19937 state is advanced by exactly 2 * n_raw words.  A draw that would make numpy's
// inversion loop restart with a fresh double (U - qn > px2: needs U within 2^-53 of 1) is reported, never guessed.
#include "common.cuh"

namespace tipb {

constexpr int ES_MT_N = 624;

__device__ __forceinline__ uint32_t es_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// keep[e] in {0,1} for draw e; status bit 1: not enough words, bit 2: an inversion restart would have been needed
__global__ void __launch_bounds__(256)
k_split_mask(const uint32_t* __restrict__ U, const uint32_t* __restrict__ pos_ptr, int64_t n_words, int64_t n_raw,
             double qn, double px2, int invert, int* __restrict__ keep, int* __restrict__ status) {
    const int64_t p0 = *pos_ptr;       // index in U of the next unread word (numpy's `pos`, 624 = block exhausted)
    if (p0 + 2 * n_raw > n_words) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, 1);
        return;
    }
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_raw) return;
    const uint2 w = make_uint2(U[p0 + 2 * e], U[p0 + 2 * e + 1]);
    const uint32_t a = es_temper(w.x) >> 5, b = es_temper(w.y) >> 6;
    const double u = (double(a) * 67108864.0 + double(b)) / 9007199254740992.0;
    const int x = u > qn;
    if (x && (u - qn) > px2) atomicOr(status, 2);
    keep[e] = invert ? 1 - x : x;
}

// a failed call consumes nothing; otherwise the state becomes (block holding the next unread word, position in it)
__global__ void __launch_bounds__(256)
k_split_advance(const uint32_t* __restrict__ U, int64_t n_raw, const int* __restrict__ status, uint32_t* __restrict__ state) {
    __shared__ int64_t s_flat;
    if (threadIdx.x == 0) s_flat = (*status == 0 && n_raw > 0) ? int64_t(state[ES_MT_N]) + 2 * n_raw : -1;
    __syncthreads();
    const int64_t flat = s_flat;
    if (flat < 0) return;
    const int64_t block = (flat - 1) / ES_MT_N;
    uint32_t v[3];
    int k = 0;
    for (int i = threadIdx.x; i < ES_MT_N; i += blockDim.x) v[k++] = U[block * ES_MT_N + i];
    __syncthreads();
    k = 0;
    for (int i = threadIdx.x; i < ES_MT_N; i += blockDim.x) state[i] = v[k++];
    if (threadIdx.x == 0) state[ES_MT_N] = uint32_t(flat - block * ES_MT_N);
}

// one thread per raw pair: its relation by binary search in raw_ptr, its rank among the kept / dropped pairs of that
// relation from the scan; writes the pair and its mirror image, the relation label, and (first pair of a relation and
// empty relations via the per-relation loop below) the range rows
__global__ void __launch_bounds__(256)
k_split_emit(const int64_t* __restrict__ raw_index, const int64_t* __restrict__ raw_ptr, int64_t n_raw, int n_rel,
             const int* __restrict__ kept_scan, int64_t n_train_pairs, int64_t* __restrict__ train_idx,
             int64_t* __restrict__ train_et, int64_t* __restrict__ test_idx, int64_t* __restrict__ test_et) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_raw) return;
    int lo = 0, hi = n_rel - 1;
    while (lo < hi) {                               // last relation whose first pair is <= e (skips empty relations)
        const int mid = (lo + hi + 1) >> 1;
        if (raw_ptr[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const int64_t r0 = raw_ptr[lo], r1 = raw_ptr[lo + 1];
    const int64_t k0 = kept_scan[r0], k1 = kept_scan[r1], ke = kept_scan[e];
    const int kept = kept_scan[e + 1] - int(ke);
    const int64_t a = raw_index[e], b = raw_index[n_raw + e];
    const int64_t n_test_pairs = n_raw - n_train_pairs;
    if (kept) {
        const int64_t n_r = k1 - k0, at = 2 * k0 + (ke - k0), E2 = 2 * n_train_pairs;
        train_idx[at] = a;            train_idx[E2 + at] = b;
        train_idx[at + n_r] = b;      train_idx[E2 + at + n_r] = a;
        train_et[at] = lo;            train_et[at + n_r] = lo;
    } else {
        const int64_t d0 = r0 - k0, n_r = (r1 - r0) - (k1 - k0), at = 2 * d0 + ((e - r0) - (ke - k0)), E2 = 2 * n_test_pairs;
        test_idx[at] = a;             test_idx[E2 + at] = b;
        test_idx[at + n_r] = b;       test_idx[E2 + at + n_r] = a;
        test_et[at] = lo;             test_et[at + n_r] = lo;
    }
}

__global__ void k_split_ranges(const int64_t* __restrict__ raw_ptr, int n_rel, const int* __restrict__ kept_scan,
                               int64_t* __restrict__ train_range, int64_t* __restrict__ test_range) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rel) return;
    const int64_t r0 = raw_ptr[r], r1 = raw_ptr[r + 1];
    const int64_t k0 = kept_scan[r0], k1 = kept_scan[r1];
    train_range[2 * r] = 2 * k0;
    train_range[2 * r + 1] = 2 * k1;
    test_range[2 * r] = 2 * (r0 - k0);
    test_range[2 * r + 1] = 2 * (r1 - k1);
}

}  // namespace tipb

using namespace tipb;

extern "C" {

size_t tipb_edge_split_workspace_bytes(int64_t n_raw) { return scan_ws_bytes(n_raw + 2) + 512; }

int tipb_edge_split_mask(uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words, int64_t n_raw, double qn,
                         double px2, int invert, int32_t* kept_scan, int32_t* status, void* ws, size_t ws_bytes,
                         void* stream) {
    TIPB_CHECK_ARG(mt_state && stream_words && kept_scan && status && ws, "edge_split_mask: NULL argument");
    TIPB_CHECK_ARG(n_raw >= 0 && n_raw < (int64_t(1) << 30), "edge_split_mask: bad pair count");
    TIPB_CHECK_ARG(n_words > ES_MT_N && n_words < (int64_t(1) << 32), "edge_split_mask: bad stream length");
    TIPB_CHECK_ARG(qn > 0.0 && qn <= 1.0, "edge_split_mask: qn must be exp(log(1 - min(p, 1 - p))) with 0 < p < 1");
    TIPB_CHECK_ARG(ws_bytes >= tipb_edge_split_workspace_bytes(n_raw), "edge_split_mask: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    TIPB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    TIPB_CHECK_CUDA(cudaMemsetAsync(kept_scan, 0, size_t(n_raw + 1) * sizeof(int32_t), s));
    if (n_raw > 0) {
        k_split_mask<<<(unsigned)ceil_div(n_raw, 256), 256, 0, s>>>(stream_words, mt_state + ES_MT_N, n_words, n_raw, qn, px2,
                                                                    invert, kept_scan, status);
        if (int rc = exclusive_scan_i32(kept_scan, kept_scan, n_raw, ws, s)) return rc;
    }
    k_split_advance<<<1, 256, 0, s>>>(stream_words, n_raw, status, mt_state);
    TIPB_CHECK_LAUNCH("edge_split_mask");
    return TIPB_OK;
}

int tipb_edge_split_emit(const int64_t* raw_index, const int64_t* raw_ptr, int64_t n_raw, int64_t n_rel,
                         const int32_t* kept_scan, int64_t n_train_pairs, int64_t* train_idx, int64_t* train_et,
                         int64_t* train_range, int64_t* test_idx, int64_t* test_et, int64_t* test_range, void* stream) {
    TIPB_CHECK_ARG(raw_ptr && kept_scan && train_range && test_range && n_rel > 0 && n_rel < (int64_t(1) << 30),
                   "edge_split_emit: bad argument");
    TIPB_CHECK_ARG(n_train_pairs >= 0 && n_train_pairs <= n_raw, "edge_split_emit: bad kept count");
    TIPB_CHECK_ARG(n_raw == 0 || raw_index, "edge_split_emit: NULL argument");
    TIPB_CHECK_ARG(n_train_pairs == 0 || (train_idx && train_et), "edge_split_emit: NULL train output");
    TIPB_CHECK_ARG(n_train_pairs == n_raw || (test_idx && test_et), "edge_split_emit: NULL test output");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_raw > 0)
        k_split_emit<<<(unsigned)ceil_div(n_raw, 256), 256, 0, s>>>(raw_index, raw_ptr, n_raw, (int)n_rel, kept_scan,
                                                                    n_train_pairs, train_idx, train_et, test_idx, test_et);
    k_split_ranges<<<(unsigned)ceil_div(n_rel, 256), 256, 0, s>>>(raw_ptr, (int)n_rel, kept_scan, train_range, test_range);
    TIPB_CHECK_LAUNCH("edge_split_emit");
    return TIPB_OK;
}
}
