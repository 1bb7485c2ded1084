// Shared device/host helpers for libtipb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tipb200.h"

namespace tipb {

// ---- error plumbing: nothing throws across the C ABI -------------------------------------
void set_last_error(const char* fmt, ...);

#define TIPB_CHECK_ARG(cond, ...)                      \
    do {                                               \
        if (!(cond)) {                                 \
            tipb::set_last_error(__VA_ARGS__);         \
            return TIPB_ERR_INVALID_ARGUMENT;          \
        }                                              \
    } while (0)

#define TIPB_CHECK_LAUNCH(name)                                                           \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            tipb::set_last_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return TIPB_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define TIPB_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            tipb::set_last_error("%s failed: %s", #expr, cudaGetErrorString(e__));         \
            (void)cudaGetLastError(); /* do not leave it for the next launch check */      \
            return TIPB_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

// ---- device info (cached per process; B200 = 148 SMs, 227 KB opt-in smem per CTA) ---------
int sm_count();
int max_smem_optin();
// raises the dynamic shared memory limit of a kernel once (cached per function; no-op <= 48 KB)
int ensure_dyn_smem(const void* func, size_t bytes);

// ---- host-side workspace carving, 256-byte aligned -----------------------------------------
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
    size_t used() const { return (off + 255) & ~size_t(255); }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// float4 arithmetic on Blackwell's packed FP32 pipe: FADD2 / FFMA2 / FMUL2 process two fp32 lanes per
// instruction (sm_100 only), halving the issue slots of the gather-accumulate loops, which are issue bound.
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
    const float2 ss = make_float2(s, s);
    const float2 lo = __ffma2_rn(ss, make_float2(a.x, a.y), make_float2(c.x, c.y));
    const float2 hi = __ffma2_rn(ss, make_float2(a.z, a.w), make_float2(c.z, c.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) {
    const float2 lo = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fmul2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// a.b over four lanes: two packed ops + one scalar add
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
    float2 p = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    p = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), p);
    return p.x + p.y;
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// streaming (read-once) global loads that do not pollute L1
__device__ __forceinline__ int ld_stream_i32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#endif

// ---- typed CSR view (pointers into one caller-owned plan buffer; typed_csr.cu) --------------
struct CsrView {
    int64_t entries, n_nodes, n_rel, seg_cap;
    int* counts;
    int* eid;
    int* other;
    int* seg_ptr;
    int* seg_node;
    int* seg_rel;
    int* node_ptr;
    int* deg;
    float* inv_deg;
    int* rel_seg_ptr;
    int* rel_seg;
};
CsrView csr_view(const void* plan, int64_t entries, int64_t n_nodes, int64_t n_rel);
size_t csr_build_ws_bytes(int64_t entries, int64_t n_nodes, int64_t n_rel);
int csr_build(const int64_t* edge_index, const int64_t* edge_type, const int64_t* range_list, int64_t E,
              int64_t n_nodes, int64_t n_other, int64_t n_rel, int by_src, int doubled, int drop_loops, int rel_major,
              void* plan, void* ws, cudaStream_t s);

// ---- internal cross-file entry points (all enqueue on `s`, never synchronise) ---------------
size_t scan_ws_bytes(int64_t n);
// exclusive scan of in[0..n) into out[0..n); out[n] (one past) receives the total.  in may alias out.
int exclusive_scan_i32(const int* in, int* out, int64_t n, void* ws, cudaStream_t s);

// tensor-core decoder sweep (sweep_tc.cu)
bool sweep_tc_supported(int64_t n_nodes, int64_t n_rel, int dim);
int sweep_tc_run(const float* z, const float* w, int64_t n_nodes, int64_t n_rel, int dim, int apply_sigmoid, float* out,
                 cudaStream_t s);
int* sweep_tc_error_flag();

size_t sort_ws_bytes(int64_t n);
// stable LSD radix sort of (key,val) pairs on the low `key_bits` bits of key.  Result lands in
// (keys_out, vals_out); (keys_in, vals_in) are clobbered.
int sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                   int key_bits, void* ws, cudaStream_t s);

}  // namespace tipb
