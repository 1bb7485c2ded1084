// Device-wide primitives used by the index builders: exclusive scan and a stable LSD radix sort.
// Hand-written (no CUB/Thrust): the index structures they produce are part of the bit-exact
// contract, so their ordering rules are spelled out here.
#include <stdarg.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace tipb {

// ------------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static int g_sm_count = 0, g_smem_optin = 0;
static void query_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
    if (g_smem_optin <= 0) g_smem_optin = 227 * 1024;
}
int sm_count() {
    if (!g_sm_count) query_device();
    return g_sm_count;
}
int max_smem_optin() {
    if (!g_smem_optin) query_device();
    return g_smem_optin;
}

int ensure_dyn_smem(const void* func, size_t bytes) {
    if (bytes <= 48 * 1024) return TIPB_OK;
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;   // the attribute is per (device, function)
    int dev = 0;
    TIPB_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = granted[std::make_pair(dev, func)];
    if (cur >= bytes) return TIPB_OK;
    // the opt-in limit covers static + dynamic shared memory of the kernel
    cudaFuncAttributes fa;
    TIPB_CHECK_CUDA(cudaFuncGetAttributes(&fa, func));
    const size_t optin = size_t(max_smem_optin());
    size_t want = optin > fa.sharedSizeBytes ? optin - fa.sharedSizeBytes : 0;
    if (bytes > want) {
        set_last_error("kernel needs %zu bytes of shared memory, device allows %zu", bytes, want);
        return TIPB_ERR_UNSUPPORTED;
    }
    TIPB_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    cur = want;
    return TIPB_OK;
}

// ================================================================================================
// exclusive scan (int32).  Three launches: chunk sums -> scan of sums (one CTA) -> chunk scans.
// ================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp /*[33]*/, int* total) {
    // inclusive scan inside the warp
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, x, o);
        if (lane_id() >= o) x += y;
    }
    if (lane_id() == 31) smem_warp[warp_id()] = x;
    __syncthreads();
    if (warp_id() == 0) {
        int nw = blockDim.x >> 5;
        int w = lane_id() < nw ? smem_warp[lane_id()] : 0;
        int xs = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FULL, xs, o);
            if (lane_id() >= o) xs += y;
        }
        if (lane_id() < nw) smem_warp[lane_id()] = xs - w;  // exclusive warp offsets
        if (lane_id() == 31) smem_warp[32] = xs;            // block total
    }
    __syncthreads();
    int res = x - v + smem_warp[warp_id()];
    if (total) *total = smem_warp[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chunk_sums(const int* __restrict__ in, int64_t n,
                                                                   int* __restrict__ sums) {
    __shared__ int sw[33];
    int64_t base = int64_t(blockIdx.x) * SCAN_CHUNK;
    int acc = 0;
    for (int i = threadIdx.x; i < SCAN_CHUNK; i += SCAN_THREADS) {
        int64_t idx = base + i;
        if (idx < n) acc += in[idx];
    }
    acc = warp_sum_i(acc);
    if (lane_id() == 0) sw[warp_id()] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += sw[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_sums(int* __restrict__ sums, int64_t nb, int* __restrict__ total_out) {
    __shared__ int sw[33];
    int carry = 0;
    for (int64_t base = 0; base < nb; base += 1024) {
        int64_t idx = base + threadIdx.x;
        int v = idx < nb ? sums[idx] : 0;
        int tot;
        int ex = block_exclusive_scan(v, sw, &tot);
        if (idx < nb) sums[idx] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chunks(const int* in, int* out, int64_t n,
                                                               const int* __restrict__ sums) {
    __shared__ int sw[33];
    int64_t base = int64_t(blockIdx.x) * SCAN_CHUNK + int64_t(threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int local = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        local += v[i];
    }
    int off = block_exclusive_scan(local, sw, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = off;
        off += v[i];
    }
}

// Single-pass form (decoupled look-back): chunks are claimed in ticket order; a chunk publishes its aggregate,
// then its warp 0 walks the published words of its predecessors 32 at a time until it meets an inclusive prefix,
// and publishes its own.  Flag and value travel in ONE 64-bit word, so no fence is needed between them.
//   state[c] = (flag << 32) | value,   flag 0: nothing yet, 1: chunk aggregate, 2: inclusive prefix
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(const int* in, int* out, int64_t n, unsigned long long* __restrict__ state, int* __restrict__ ticket) {
    __shared__ int sw[33];
    __shared__ int s_chunk, s_excl;
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1);
    __syncthreads();
    const int chunk = s_chunk;
    const int64_t base = int64_t(chunk) * SCAN_CHUNK + int64_t(threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int local = 0;
    if (base + SCAN_ITEMS <= n) {
        const int4* p = reinterpret_cast<const int4*>(in + base);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS / 4; ++i) {
            const int4 q = p[i];
            v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = base + i < n ? in[base + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) local += v[i];
    int total;
    const int ex = block_exclusive_scan(local, sw, &total);
    if (warp_id() == 0) {
        const int lane = lane_id();
        volatile unsigned long long* vs = state;
        int excl = 0;
        if (chunk > 0) {
            if (lane == 0) atomicExch(&state[chunk], (1ull << 32) | (unsigned)total);
            int look = chunk - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long st = idx >= 0 ? vs[idx] : (2ull << 32);
                while (__any_sync(FULL, (st >> 32) == 0)) {
                    if ((st >> 32) == 0) st = vs[idx];
                }
                const unsigned pm = __ballot_sync(FULL, (st >> 32) == 2);
                const int first = pm ? __ffs(pm) - 1 : 31;   // nearest predecessor that already holds a prefix
                int val = lane <= first ? int(unsigned(st)) : 0;
                val = warp_sum_i(val);
                excl += val;
                if (pm) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            atomicExch(&state[chunk], (2ull << 32) | (unsigned)(excl + total));
            s_excl = excl;
        }
    }
    __syncthreads();
    int off = ex + s_excl;
    if (base + SCAN_ITEMS <= n) {
        int4* p = reinterpret_cast<int4*>(out + base);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS / 4; ++i) {
            int4 q;
            q.x = off; off += v[4 * i];
            q.y = off; off += v[4 * i + 1];
            q.z = off; off += v[4 * i + 2];
            q.w = off; off += v[4 * i + 3];
            p[i] = q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            if (base + i < n) out[base + i] = off;
            off += v[i];
        }
    }
    // one past the end receives the total
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = s_excl + ex + local;
}

size_t scan_ws_bytes(int64_t n) { return size_t(ceil_div(n > 0 ? n : 1, SCAN_CHUNK) + 2) * sizeof(unsigned long long) + 256; }

int exclusive_scan_i32(const int* in, int* out, int64_t n, void* ws, cudaStream_t s) {
    if (n <= 0) {
        TIPB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int), s));
        return TIPB_OK;
    }
    const int64_t nb = ceil_div(n, SCAN_CHUNK);
    unsigned long long* state = static_cast<unsigned long long*>(ws);
    int* ticket = reinterpret_cast<int*>(state + nb);
    TIPB_CHECK_CUDA(cudaMemsetAsync(ws, 0, size_t(nb + 1) * sizeof(unsigned long long), s));
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) {
        // unaligned views: the three-launch form has no vector accesses
        int* sums = reinterpret_cast<int*>(state);
        k_scan_chunk_sums<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, sums);
        k_scan_sums<<<1, 1024, 0, s>>>(sums, nb, out + n);
        k_scan_chunks<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, out, n, sums);
    } else {
        k_scan_lookback<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, out, n, state, ticket);
    }
    TIPB_CHECK_LAUNCH("exclusive_scan_i32");
    return TIPB_OK;
}

// ================================================================================================
// stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits per pass.
//   tile = 8 warps x 512 consecutive elements; element order inside a tile is
//   (warp, round, lane), which equals memory order, so equal digits keep their input order.
// ================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_PER_WARP = 32 * RS_ROUNDS;
constexpr int RS_TILE = RS_WARPS * RS_PER_WARP;
constexpr int RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                        int* __restrict__ hist, int num_tiles) {
    __shared__ int h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = int64_t(blockIdx.x) * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        int64_t idx = base + i;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (RADIX - 1)], 1);
    }
    __syncthreads();
    hist[int64_t(threadIdx.x) * num_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys_in,
                                                           const uint32_t* __restrict__ vals_in,
                                                           uint32_t* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                           const int* __restrict__ hist_scanned, int num_tiles) {
    __shared__ int wh[RS_WARPS][RADIX];
    __shared__ int gb[RADIX];
    const int w = warp_id(), lane = lane_id();
    for (int i = threadIdx.x; i < RS_WARPS * RADIX; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();

    const int64_t wbase = int64_t(blockIdx.x) * RS_TILE + int64_t(w) * RS_PER_WARP;
    // pass 1: per-warp digit counts
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        unsigned act = __ballot_sync(FULL, valid);
        if (valid) {
            int d = (keys_in[idx] >> shift) & (RADIX - 1);
            unsigned peers = __match_any_sync(act, d);
            if ((__ffs(peers) - 1) == lane) wh[w][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps, per digit; global base of this tile's run of that digit
    {
        int d = threadIdx.x;
        int run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            int c = wh[ww][d];
            wh[ww][d] = run;
            run += c;
        }
        gb[d] = hist_scanned[int64_t(d) * num_tiles + blockIdx.x];
    }
    __syncthreads();
    // pass 2: stable placement
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        unsigned act = __ballot_sync(FULL, valid);
        uint32_t k = 0, v = 0;
        int d = 0, pos = 0;
        unsigned peers = 0;
        if (valid) {
            k = keys_in[idx];
            v = vals_in[idx];
            d = (k >> shift) & (RADIX - 1);
            peers = __match_any_sync(act, d);
            pos = gb[d] + wh[w][d] + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
        if (valid && (__ffs(peers) - 1) == lane) wh[w][d] += __popc(peers);
        __syncwarp();
        if (valid) {
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
    }
}

size_t sort_ws_bytes(int64_t n) {
    int64_t tiles = ceil_div(n > 0 ? n : 1, RS_TILE);
    size_t hist = size_t(RADIX) * tiles + 1;
    return ((hist * sizeof(int) + 255) & ~size_t(255)) + scan_ws_bytes(int64_t(hist)) + 256;
}

int sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                   int key_bits, void* ws, cudaStream_t s) {
    if (key_bits < 1) key_bits = 1;
    int passes = (key_bits + 7) / 8;
    if (n <= 0) return TIPB_OK;
    int64_t tiles = ceil_div(n, RS_TILE);
    Carver c(ws);
    int* hist = c.take<int>(size_t(RADIX) * tiles + 1);
    void* scan_ws = c.take<char>(scan_ws_bytes(RADIX * tiles));
    // an even pass count would end in the input buffers: prepend a copy pass by sorting on a
    // zero-width digit is wasteful, so instead run one extra (stable, harmless) pass on the top bits.
    if ((passes & 1) == 0) passes += 1;
    uint32_t *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
    for (int p = 0; p < passes; ++p) {
        int shift = p * 8;
        if (shift >= 32) shift = 24;  // idempotent extra pass (already sorted on these bits)
        k_rs_hist<<<(unsigned)tiles, RS_THREADS, 0, s>>>(ki, n, shift, hist, (int)tiles);
        int rc = exclusive_scan_i32(hist, hist, RADIX * tiles, scan_ws, s);
        if (rc) return rc;
        k_rs_scatter<<<(unsigned)tiles, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, hist, (int)tiles);
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    TIPB_CHECK_LAUNCH("sort_pairs_u32");
    return TIPB_OK;
}

}  // namespace tipb

// ------------------------------------------------------------------------------------------------
extern "C" {

int tipb_version(void) { return TIPB_VERSION; }
const char* tipb_last_error(void) { return tipb::g_last_error; }

size_t tipb_sort_workspace_bytes(int64_t n) { return tipb::sort_ws_bytes(n); }

int tipb_sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                        int key_bits, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "sort: n=%lld out of range", (long long)n);
    TIPB_CHECK_ARG(ws_bytes >= tipb::sort_ws_bytes(n), "sort: workspace too small");
    return tipb::sort_pairs_u32(keys_in, vals_in, keys_out, vals_out, n, key_bits, ws, (cudaStream_t)stream);
}

size_t tipb_scan_workspace_bytes(int64_t n) { return tipb::scan_ws_bytes(n); }

int tipb_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, void* stream) {
    TIPB_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "scan: n=%lld out of range", (long long)n);
    TIPB_CHECK_ARG(ws_bytes >= tipb::scan_ws_bytes(n), "scan: workspace too small");
    return tipb::exclusive_scan_i32(in, out, n, ws, (cudaStream_t)stream);
}
}
