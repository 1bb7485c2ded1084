// Decoder sweep on the 5th-generation tensor cores (BASELINE.json config 5: all N x N drug pairs x R relations;
// reference: MultiInnerProductDecoder.forward evaluated on every pair, src/layers.py:590-592).
//
//   out[r, i, j] = act( sum_k z[i,k] w[r,k] z[j,k] )            1.43 GB of fp32 scores at 645 x 645 x 861
//
// The CUDA-core kernel (decoder.cu: k_decoder_sweep_tiled) is bound by 11.5 GFLOP of fp32 FMA plus the sigmoid
// (23 % of the HBM write roofline).  Here the contraction runs as ONE GEMM on tcgen05:
//
//   D[j, c] = sum_kk A'[j, kk] B'[c, kk],      c = r * N + i  (555,345 rows),   kk = 0 .. 6 * dim
//
// fp32 accuracy from bf16 operands: x = hi + mid + lo (three bf16 pieces, 24 significant bits) and the six products
// hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi are laid side by side along K (the dropped terms are <= 2^-24
// relative); A' holds the pieces of z[j], B' those of u = z[i] * w[r].  Accumulation is fp32 in TMEM.
//
// One persistent CTA per SM (640 threads).  A' (the pieces of all of z, 6 M-tiles of 128 rows) is built once; per
// N-tile of 256 rows c:
//   builder warps   compute u, split it, store B' (double buffered) in the canonical K-major no-swizzle UMMA layout,
//                   fence.proxy.async, arrive
//   MMA thread      per M-tile: wait for a free TMEM accumulator (2 x 256 columns), issue the six piece products as
//                   tcgen05.mma (M 128, N 256, K 16) selecting the pieces through the descriptor start addresses,
//                   tcgen05.commit to the "full" mbarrier
//   epilogue warps  (16: four per TMEM lane quarter) request both of their 32-column chunks from TMEM before waiting
//                   (one tcgen05.ld at a time reads TMEM at 30 B/clk/SM, two in flight on 16 warps at > 120), apply
//                   the sigmoid (ex2 + rcp), stage the chunk in shared memory and write it as LINE-ALIGNED stores: the
//                   output is one flat array whose rows (645 floats) start at arbitrary 4-byte offsets, and warp
//                   stores that straddle lines run at ~60 % of the rate of whole lines (measured: 323 vs 500 us)
// Every barrier wait is bounded (a protocol fault sets a flag, tipb_decoder_sweep_status, instead of hanging the GPU).
// Measured on B200 (tools/ubench_sweep.py, profiles/): 415 us = 3.45 TB/s written = 53 % of the measured copy peak, from
// 931 us for the CUDA-core kernel; MMAs + handshakes alone 59 us, + TMEM loads 69 us, + sigmoid and staging 242 us:
// the rest is the store loop (two partial lines per 512-byte row segment).  Bits 1..6 of `apply_sigmoid` switch parts of
// the epilogue off -- they exist for exactly these measurements and are not part of the ABI contract.
#include "umma.cuh"

namespace tipb {

constexpr int ST_THREADS = 640;          // warps 0-15: epilogue, warp 16: MMA issue, warps 17-19: B' builders
constexpr int ST_EPI_WARPS = 16;         // warp w drains TMEM lanes 32 (w % 4) .. + 32; the four warps w / 4 = g form group g
constexpr int ST_GROUPS = ST_EPI_WARPS / 4;
constexpr int ST_BUILD_THREADS = ST_THREADS - 32 * (ST_EPI_WARPS + 1);
constexpr int ST_TILE_N = 256;           // rows c per N-tile = MMA N = TMEM columns per accumulator
constexpr int ST_TILE_M = 128;           // rows j per M-tile = MMA M = TMEM lanes
constexpr int ST_MAX_MTILES = 6;         // n_nodes <= 768
constexpr int ST_CH_PER_WARP = ST_TILE_N / 32 / ST_GROUPS;     // 32-column chunks of an accumulator per warp: 2

// operand row layout: the three bf16 pieces [hi | mid | lo] of a row side by side, DIM = 16 elements (one K = 16 MMA
// step) each; the six products pick their pieces through the descriptor start address
template <int DIM>
struct SweepLayout {
    static_assert(DIM == 16, "one piece = one K = 16 step");
    static constexpr int K = 3 * DIM;                 // bf16 elements per row
    static constexpr int CHUNKS = K / 8;              // 16-byte chunks per row (6)
    static constexpr int LBO = 128;                   // adjacent core matrices along K
    static constexpr int SBO = CHUNKS * 128;          // next 8-row group
    static constexpr int PIECE_BYTES = 2 * LBO;       // one piece = two core matrices along K
    static constexpr int A_TILE_BYTES = (ST_TILE_M / 8) * SBO;
    static constexpr int B_TILE_BYTES = (ST_TILE_N / 8) * SBO;
    __device__ static uint32_t chunk_offset(int row, int chunk) { return uint32_t((row >> 3) * SBO + chunk * 128 + (row & 7) * 16); }
};
// products hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi (piece index 0 = hi, 1 = mid, 2 = lo)
__device__ __constant__ int c_piece_a[6] = {0, 0, 1, 1, 0, 2};
__device__ __constant__ int c_piece_b[6] = {0, 1, 0, 1, 2, 0};

constexpr int ST_STAGE_ROWS = 32, ST_STAGE_PITCH = ST_TILE_M + 4;      // staging of one 32-column chunk: [32 c][128 j]

template <int DIM>
__global__ void __launch_bounds__(ST_THREADS, 1)
k_decoder_sweep_tc(const float* __restrict__ z, const float* __restrict__ w, int n_nodes, int n_rel, int apply_sigmoid,
                   float* __restrict__ out, int* __restrict__ error_flag) {
    using LY = SweepLayout<DIM>;
    extern __shared__ __align__(1024) uint8_t st_smem[];
    const int m_tiles = (n_nodes + ST_TILE_M - 1) / ST_TILE_M;
    uint8_t* sA = st_smem;                                           // [m_tiles] A' tiles
    uint8_t* sB = sA + size_t(m_tiles) * LY::A_TILE_BYTES;            // two B' tiles (double buffered)
    float* sStage = reinterpret_cast<float*>(sB + 2 * LY::B_TILE_BYTES);   // [groups][32][PITCH] epilogue staging
    __shared__ uint64_t bar_full[2], bar_empty[2], bar_b_ready[2], bar_b_free[2];
    __shared__ uint32_t s_tmem_base;

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t n_rows = int64_t(n_rel) * n_nodes;                 // rows c of B'
    const int n_tiles = int((n_rows + ST_TILE_N - 1) / ST_TILE_N);

    // ---- setup: barriers, TMEM, A'
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_full[b], 1);
            mbar_init(&bar_empty[b], ST_EPI_WARPS);
            mbar_init(&bar_b_ready[b], ST_BUILD_THREADS);
            mbar_init(&bar_b_free[b], 1);
        }
        fence_barrier_init();
    }
    if (wid == ST_EPI_WARPS) tmem_alloc(&s_tmem_base, 512);
    // A'[j, piece, k]; rows j >= n_nodes are zero
    for (int idx = tid; idx < m_tiles * ST_TILE_M * (DIM / 8); idx += ST_THREADS) {
        const int j = idx / (DIM / 8), c8 = idx % (DIM / 8);
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = j < n_nodes ? z[size_t(j) * DIM + c8 * 8 + i] : 0.f;
        const Pieces8 p = split8(x);
        uint8_t* tile = sA + size_t(j / ST_TILE_M) * LY::A_TILE_BYTES;
        const int row = j % ST_TILE_M;
        *reinterpret_cast<uint4*>(tile + LY::chunk_offset(row, 0 * (DIM / 8) + c8)) = p.hi;
        *reinterpret_cast<uint4*>(tile + LY::chunk_offset(row, 1 * (DIM / 8) + c8)) = p.mid;
        *reinterpret_cast<uint4*>(tile + LY::chunk_offset(row, 2 * (DIM / 8) + c8)) = p.lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (wid < ST_EPI_WARPS) {
        // =========================================================== epilogue: TMEM -> sigmoid -> staging -> global
        // The output is one flat array (row c of B' = 645 consecutive floats): a warp store of 32 consecutive floats
        // starts at an arbitrary 4-byte offset, and L2 handles the two partial sectors of every such store at about
        // half the rate of whole lines.  So the four warps of a group (one per TMEM lane quarter) stage their
        // [32 c][128 j] chunk in shared memory, and every warp then writes 8 of the rows as LINE-ALIGNED stores: lane l
        // takes element 32 k + l - a of the row segment, a = the segment's offset inside its 128-byte line.
        const int quarter = wid & 3, group = wid >> 2;
        float* stage = sStage + size_t(group) * ST_STAGE_ROWS * ST_STAGE_PITCH;
        const int bar_id = 1 + group;                                 // named barrier of the group's four warps
        uint32_t it = 0;
        bool ok = true;
        for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x) {
            const int64_t c0 = int64_t(t) * ST_TILE_N;
            for (int m = 0; m < m_tiles && ok; ++m, ++it) {
                const uint32_t b = it & 1u;
                ok = mbar_wait(&bar_full[b], (it >> 1) & 1u);
                if (!ok) break;
                tc_fence_after();
                const int seg_len = min(ST_TILE_M, n_nodes - m * ST_TILE_M);     // valid j of this M-tile
                const bool have_lanes = quarter * 32 < seg_len;                  // (the last M-tile is mostly padding)
                // both chunks of this warp are requested from TMEM before the first one is waited for
                uint32_t v[ST_CH_PER_WARP][32];
                if (have_lanes && !(apply_sigmoid & 32)) {
#pragma unroll
                    for (int u = 0; u < ST_CH_PER_WARP; ++u)
                        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + b * ST_TILE_N + (group + u * ST_GROUPS) * 32, v[u]);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int u = 0; u < ST_CH_PER_WARP; ++u) {
                    const int ch = group + u * ST_GROUPS;
                    if (apply_sigmoid & 16) {                      // (measurement aid: no staging, no stores)
                        if (have_lanes && (v[u][0] ^ v[u][17]) == 0x12345678u) out[0] = 0.f;
                        continue;
                    }
                    if (have_lanes) {
                        // (measured: replacing rcp by three Newton steps on the FMA pipe -- half the SFU work, 2.5x the
                        // instructions -- made the sweep slower, 460 -> 513 us: the sixteen epilogue warps are bound by
                        // instruction issue and latency, not by the SFU)
                        if (apply_sigmoid & 1) {                   // 1 / (1 + 2^(-v log2 e)): two SFU operations per score
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                float e, r;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(v[u][q]) * -1.4426950408889634f));
                                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
                                v[u][q] = __float_as_uint(r);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 32; ++q) stage[q * ST_STAGE_PITCH + quarter * 32 + lane] = __uint_as_float(v[u][q]);
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (!(apply_sigmoid & 8)) {
                        // rows quarter*8 .. +8 of the chunk; per row one per-thread pointer, the five line stores use
                        // immediate offsets from it
                        const int64_t c_first = c0 + ch * 32 + quarter * 8;
                        int64_t f = c_first * n_nodes + m * ST_TILE_M;           // flat index of the row segment's first score
                        const int rows = int(min(int64_t(8), n_rows - c_first));
#pragma unroll 1
                        for (int rr = 0; rr < rows; ++rr, f += n_nodes) {
                            const int a = int(f & 31);
                            float* __restrict__ p = out + (f - a) + lane;       // element `lane` of the segment's first line
                            const float* __restrict__ src = stage + (quarter * 8 + rr) * ST_STAGE_PITCH - a + lane;
                            const unsigned e0 = unsigned(lane - a);             // index inside the segment (wraps below 0)
#pragma unroll
                            for (int k = 0; k < ST_TILE_M / 32 + 1; ++k) {
                                if ((apply_sigmoid & 64) && (k == 0 || k == ST_TILE_M / 32)) continue;   // (aid: whole lines only)
                                if (e0 + unsigned(k * 32) < unsigned(seg_len)) p[k * 32] = src[k * 32];
                            }
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // staging is reused by the next chunk
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[b]);      // one arrival per epilogue warp
            }
        }
        if (!ok) atomicExch(error_flag, 1);
    } else if (wid == ST_EPI_WARPS) {
        // =========================================================== MMA issue (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(ST_TILE_M, ST_TILE_N);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            uint32_t it = 0, tile_no = 0;
            bool ok = true;
            for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++tile_no) {
                const uint32_t sb = tile_no & 1u;
                ok = mbar_wait(&bar_b_ready[sb], (tile_no >> 1) & 1u);
                if (!ok) break;
                tc_fence_after();
                for (int m = 0; m < m_tiles && ok; ++m, ++it) {
                    const uint32_t b = it & 1u;
                    ok = mbar_wait(&bar_empty[b], ((it >> 1) & 1u) ^ 1u);    // passes at once the first two times
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + b * ST_TILE_N;
#pragma unroll
                    for (int p = 0; p < 6; ++p) {
                        const uint64_t da = umma_desc(a_base + m * LY::A_TILE_BYTES + c_piece_a[p] * LY::PIECE_BYTES, LY::LBO, LY::SBO);
                        const uint64_t db = umma_desc(b_base + sb * LY::B_TILE_BYTES + c_piece_b[p] * LY::PIECE_BYTES, LY::LBO, LY::SBO);
                        umma_bf16(d_tmem, da, db, idesc, p > 0 ? 1u : 0u);
                    }
                    umma_commit(&bar_full[b]);          // arrives when these MMAs have written the accumulator
                }
                umma_commit(&bar_b_free[sb]);           // ... and when every MMA that read this B' tile is done
            }
            if (!ok) atomicExch(error_flag, 2);
        }
    } else {
        // =========================================================== B' builders (two tiles ahead of the epilogue)
        const int bt = tid - 32 * (ST_EPI_WARPS + 1);
        uint32_t tile_no = 0;
        bool ok = true;
        for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++tile_no) {
            const uint32_t sb = tile_no & 1u;
            if (tile_no >= 2) {
                ok = mbar_wait(&bar_b_free[sb], ((tile_no >> 1) - 1) & 1u);
                if (!ok) break;
            }
            uint8_t* tileB = sB + size_t(sb) * LY::B_TILE_BYTES;
            const int64_t c0 = int64_t(t) * ST_TILE_N;
            for (int idx = bt; idx < ST_TILE_N * (DIM / 8); idx += ST_BUILD_THREADS) {
                const int row = idx / (DIM / 8), c8 = idx % (DIM / 8);
                const int64_t c = c0 + row;
                float x[8];
                if (c < n_rows) {
                    const int r = int(uint32_t(c) / uint32_t(n_nodes)), i = int(uint32_t(c) - uint32_t(r) * uint32_t(n_nodes));
                    const float4 z0 = *reinterpret_cast<const float4*>(z + size_t(i) * DIM + c8 * 8);
                    const float4 z1 = *reinterpret_cast<const float4*>(z + size_t(i) * DIM + c8 * 8 + 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(w + size_t(r) * DIM + c8 * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(w + size_t(r) * DIM + c8 * 8 + 4);
                    x[0] = z0.x * w0.x; x[1] = z0.y * w0.y; x[2] = z0.z * w0.z; x[3] = z0.w * w0.w;
                    x[4] = z1.x * w1.x; x[5] = z1.y * w1.y; x[6] = z1.z * w1.z; x[7] = z1.w * w1.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = 0.f;
                }
                const Pieces8 p = split8(x);
                *reinterpret_cast<uint4*>(tileB + LY::chunk_offset(row, 0 * (DIM / 8) + c8)) = p.hi;
                *reinterpret_cast<uint4*>(tileB + LY::chunk_offset(row, 1 * (DIM / 8) + c8)) = p.mid;
                *reinterpret_cast<uint4*>(tileB + LY::chunk_offset(row, 2 * (DIM / 8) + c8)) = p.lo;
            }
            fence_proxy_async();                        // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&bar_b_ready[sb]);
        }
        if (!ok) atomicExch(error_flag, 3);
    }

    tc_fence_before();
    __syncthreads();
    if (wid == ST_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static size_t sweep_tc_smem(int64_t n_nodes, int dim) {
    const size_t sbo = size_t(3 * dim / 8) * 128;
    return size_t(ceil_div(n_nodes, ST_TILE_M)) * (ST_TILE_M / 8) * sbo + 2 * size_t(ST_TILE_N / 8) * sbo +
           size_t(ST_EPI_WARPS / 4) * ST_STAGE_ROWS * ST_STAGE_PITCH * sizeof(float) + 1024;
}

template <int DIM>
static int sweep_tc_launch(const float* z, const float* w, int64_t n_nodes, int64_t n_rel, int apply_sigmoid, float* out,
                           int* error_flag, cudaStream_t s) {
    const size_t smem = sweep_tc_smem(n_nodes, DIM);
    auto kern = k_decoder_sweep_tc<DIM>;
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    const int64_t n_tiles = ceil_div(n_rel * n_nodes, ST_TILE_N);
    const int grid = int(n_tiles < sm_count() ? n_tiles : sm_count());
    kern<<<grid, ST_THREADS, smem, s>>>(z, w, (int)n_nodes, (int)n_rel, apply_sigmoid, out, error_flag);
    TIPB_CHECK_LAUNCH("decoder_sweep_tc");
    return TIPB_OK;
}

__device__ int g_sweep_tc_error = 0;      // set by a role whose bounded barrier wait timed out (never in a correct run)

int* sweep_tc_error_flag() {
    static int* ptr = nullptr;
    if (!ptr) cudaGetSymbolAddress(reinterpret_cast<void**>(&ptr), g_sweep_tc_error);
    return ptr;
}

// 1 if the tensor-core sweep handles this shape (the kernels of decoder.cu take the rest)
bool sweep_tc_supported(int64_t n_nodes, int64_t n_rel, int dim) {
    if (dim != 16) return false;
    if (n_nodes < 1 || n_nodes > int64_t(ST_MAX_MTILES) * ST_TILE_M || n_rel < 1) return false;
    return sweep_tc_smem(n_nodes, dim) + 1024 <= size_t(max_smem_optin()) && n_rel * n_nodes < (int64_t(1) << 31);
}

int sweep_tc_run(const float* z, const float* w, int64_t n_nodes, int64_t n_rel, int dim, int apply_sigmoid, float* out,
                 cudaStream_t s) {
    (void)dim;
    return sweep_tc_launch<16>(z, w, n_nodes, n_rel, apply_sigmoid, out, sweep_tc_error_flag(), s);
}

}  // namespace tipb
