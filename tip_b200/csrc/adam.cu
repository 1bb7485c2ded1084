// Multi-tensor Adam: the optimiser step of tip.py:21-30,
//     optimizer = torch.optim.Adam(model.parameters(), lr=settings.lr);  optimizer.step()
// for all parameter tensors of the model (13 for TIP) in ONE launch.  Same update as torch.optim.Adam with its
// defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay, no amsgrad):
//     m += (g - m)(1 - b1);   v = b2 v + (1 - b2) g g;   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step counter lives on the device and is advanced by the kernel, so the launch is CUDA-graph capturable; the
// tensor table travels in the kernel arguments (no host->device copy, no device-side pointer table).
#include "common.cuh"

namespace tipb {

constexpr int ADAM_MAX_TENSORS = 48;
constexpr int ADAM_CHUNK = 4096;  // elements per CTA trip

struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    int chunk_begin[ADAM_MAX_TENSORS + 1];  // prefix of ceil(n / ADAM_CHUNK)
    int64_t n[ADAM_MAX_TENSORS];
    int count;
};

__global__ void __launch_bounds__(256)
k_adam(const AdamTable tab, float lr, float beta1, float beta2, float omb1, float omb2, float eps,
       const float* __restrict__ step_in) {
    // step_in holds t-1; every CTA derives the same bias corrections (the counter itself is advanced by k_adam_tick)
    // (bias corrections in double, as torch derives them from Python floats)
    const double t = double(step_in[0]) + 1.0;
    const double bc1 = 1.0 - pow(1.0 - double(omb1), t), bc2 = 1.0 - pow(1.0 - double(omb2), t);
    const float step_size = float(double(lr) / bc1), rb2 = float(sqrt(bc2));
    int lo = 0, hi = tab.count - 1;
    const int c = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab.chunk_begin[mid] <= c) lo = mid; else hi = mid - 1;
    }
    const int64_t base = int64_t(c - tab.chunk_begin[lo]) * ADAM_CHUNK;
    const int64_t n = tab.n[lo];
    float* __restrict__ p = tab.p[lo];
    const float* __restrict__ g = tab.g[lo];
    float* __restrict__ m = tab.m[lo];
    float* __restrict__ v = tab.v[lo];
    for (int64_t i = base + threadIdx.x; i < base + ADAM_CHUNK && i < n; i += 256) {
        const float gi = g[i];
        float mi = m[i], vi = v[i];
        mi = mi + (gi - mi) * omb1;              // lerp(m, g, 1 - beta1)
        vi = vi * beta2 + omb2 * (gi * gi);      // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / rb2 + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

__global__ void k_adam_tick(float* step) { step[0] += 1.0f; }

}  // namespace tipb

using namespace tipb;

extern "C" {

int tipb_adam_max_tensors(void) { return ADAM_MAX_TENSORS; }

int tipb_adam_step(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                   void* const* exp_avg_sq, const int64_t* numel, double lr_d, double beta1_d, double beta2_d, double eps_d,
                   float* step_dev, void* stream) {
    const float lr = float(lr_d), beta1 = float(beta1_d), beta2 = float(beta2_d), eps = float(eps_d);
    TIPB_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel)) && step_dev,
                   "adam_step: NULL argument");
    TIPB_CHECK_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                   "adam_step: bad hyper-parameter");
    cudaStream_t s = (cudaStream_t)stream;
    for (int first = 0; first < n_tensors; first += ADAM_MAX_TENSORS) {
        AdamTable tab;
        const int cnt = n_tensors - first < ADAM_MAX_TENSORS ? n_tensors - first : ADAM_MAX_TENSORS;
        int chunks = 0;
        int used = 0;
        for (int k = 0; k < cnt; ++k) {
            const int64_t n = numel[first + k];
            TIPB_CHECK_ARG(n >= 0 && n < (int64_t(1) << 40), "adam_step: bad tensor size");
            if (n == 0) continue;
            TIPB_CHECK_ARG(params[first + k] && grads[first + k] && exp_avg[first + k] && exp_avg_sq[first + k],
                           "adam_step: NULL tensor");
            tab.p[used] = (float*)params[first + k];
            tab.g[used] = (const float*)grads[first + k];
            tab.m[used] = (float*)exp_avg[first + k];
            tab.v[used] = (float*)exp_avg_sq[first + k];
            tab.n[used] = n;
            tab.chunk_begin[used] = chunks;
            chunks += int(ceil_div(n, ADAM_CHUNK));
            ++used;
        }
        tab.chunk_begin[used] = chunks;
        tab.count = used;
        // 1 - beta is formed in double (as torch forms it from Python floats): 1.0f - 0.999f is off by 1.3e-5
        if (chunks > 0)
            k_adam<<<chunks, 256, 0, s>>>(tab, lr, beta1, beta2, float(1.0 - double(beta1_d)), float(1.0 - double(beta2_d)), eps,
                                          step_dev);
    }
    k_adam_tick<<<1, 1, 0, s>>>(step_dev);
    TIPB_CHECK_LAUNCH("adam_step");
    return TIPB_OK;
}
}
