// SURVEY 8 f, row N2: the MLP baseline (models/mlp.py:7-27) between its Linear layers.
//   [Linear] -> BatchNorm1d -> Dropout(0.1) -> ReLU           (x is [batch][features], row-major)
// The Linear layers are dense GEMMs on gsr_tcgemm_f32 / gsr_bgemm_f32; the three layers in between are one
// memory-bound pass here (batch statistics first when training).  The dropout mask is a counter-based hash of
// (seed, element index): nothing is stored, the backward pass recomputes it, and the CPU restatement used by the
// tests reproduces it bit for bit.
#include "common.cuh"

namespace gsr {
namespace mlp {

constexpr int TX = 32, TY = 8;

// splitmix64 of seed + golden * index, top 24 bits as a uniform in [0, 1)
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long index) {
    unsigned long long z = seed + index * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// per-feature sum over the batch of f(b): threads (tx, ty) = (feature, batch phase), fixed order
template <class F>
__device__ __forceinline__ float column_sum(F f, int batch, float (*red)[TX + 1]) {
    float acc = 0.f;
    for (int b = threadIdx.y; b < batch; b += TY) acc += f(b);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    float total = 0.f;
#pragma unroll
    for (int j = 0; j < TY; ++j) total += red[j][threadIdx.x];
    __syncthreads();
    return total;
}

// nn.BatchNorm1d training statistics (models/mlp.py:14): mean, biased variance (two passes), and the running
// estimates with the unbiased variance, momentum as torch.
__global__ void __launch_bounds__(TX * TY) bn_stats_kernel(const float* __restrict__ x, int batch, int features,
                                                           float* __restrict__ mean, float* __restrict__ var,
                                                           float* running_mean, float* running_var, float momentum) {
    __shared__ float red[TY][TX + 1];
    const int h = blockIdx.x * TX + threadIdx.x;
    const bool live = h < features;
    const float* col = x + (live ? h : 0);
    const float m = column_sum([&](int b) { return live ? __ldg(col + (size_t)b * features) : 0.f; }, batch, red) / (float)batch;
    const float ss = column_sum([&](int b) { const float d = live ? __ldg(col + (size_t)b * features) - m : 0.f; return d * d; }, batch, red);
    if (live && threadIdx.y == 0) {
        mean[h] = m;
        var[h] = ss / (float)batch;
        if (running_mean) running_mean[h] = (1.f - momentum) * running_mean[h] + momentum * m;
        if (running_var) running_var[h] = (1.f - momentum) * running_var[h] + momentum * (batch > 1 ? ss / (float)(batch - 1) : ss);
    }
}

__device__ __forceinline__ float keep_scale(float p_drop, unsigned long long seed, unsigned long long index) {
    if (p_drop <= 0.f) return 1.f;
    return uniform01(seed, index) >= p_drop ? 1.f / (1.f - p_drop) : 0.f;
}

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ var, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, float p_drop,
                                                         unsigned long long seed, int relu, float* __restrict__ y,
                                                         long long total, int features) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(e % features);
        const float rstd = 1.f / sqrtf(__ldg(var + h) + eps);
        float z = (__ldg(x + e) - __ldg(mean + h)) * rstd * __ldg(gamma + h) + __ldg(beta + h);
        z *= keep_scale(p_drop, seed, (unsigned long long)e);
        y[e] = relu ? fmaxf(z, 0.f) : z;
    }
}

// g = dy through ReLU and dropout; d beta = sum_b g, d gamma = sum_b g * xhat
__device__ __forceinline__ float upstream(float dy, float xhat, float gamma, float beta, float p_drop,
                                          unsigned long long seed, unsigned long long index, int relu) {
    const float k = keep_scale(p_drop, seed, index);
    if (relu && !((xhat * gamma + beta) * k > 0.f)) return 0.f;
    return dy * k;
}

__global__ void __launch_bounds__(TX * TY) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ var, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ dy,
                                                                float eps, float p_drop, unsigned long long seed, int relu,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int batch, int features) {
    __shared__ float red[TY][TX + 1];
    const int h = blockIdx.x * TX + threadIdx.x;
    const bool live = h < features;
    const int hc = live ? h : 0;
    const float m = __ldg(mean + hc), rstd = 1.f / sqrtf(__ldg(var + hc) + eps), ga = __ldg(gamma + hc), be = __ldg(beta + hc);
    auto g_of = [&](int b, float& xhat) {
        const size_t e = (size_t)b * features + hc;
        xhat = (__ldg(x + e) - m) * rstd;
        return live ? upstream(__ldg(dy + e), xhat, ga, be, p_drop, seed, e, relu) : 0.f;
    };
    const float sb = column_sum([&](int b) { float xh; return g_of(b, xh); }, batch, red);
    const float sg = column_sum([&](int b) { float xh; const float g = g_of(b, xh); return g * xh; }, batch, red);
    if (live && threadIdx.y == 0) { dbeta[h] = sb; dgamma[h] = sg; }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ var, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ dy,
                                                           const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                           float eps, float p_drop, unsigned long long seed, int relu,
                                                           int batch_stats, float* __restrict__ dx, long long total,
                                                           int batch, int features) {
    const float inv_b = 1.f / (float)batch;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(e % features);
        const float rstd = 1.f / sqrtf(__ldg(var + h) + eps), ga = __ldg(gamma + h);
        const float xhat = (__ldg(x + e) - __ldg(mean + h)) * rstd;
        float g = upstream(__ldg(dy + e), xhat, ga, __ldg(beta + h), p_drop, seed, (unsigned long long)e, relu);
        if (batch_stats) g -= (__ldg(dbeta + h) + xhat * __ldg(dgamma + h)) * inv_b;     // the statistics depend on x
        dx[e] = ga * rstd * g;
    }
}

static int grid_for(long long total) { return (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 8); }

}  // namespace mlp
}  // namespace gsr

using namespace gsr;

extern "C" int gsr_bn_stats_f32(const float* x, int batch, int features, float* mean, float* var,
                                float* running_mean, float* running_var, float momentum, void* stream) {
    if (!x || !mean || !var || batch <= 0 || features <= 0) return fail(GSR_E_ARG, "bn_stats: bad argument");
    mlp::bn_stats_kernel<<<(features + mlp::TX - 1) / mlp::TX, dim3(mlp::TX, mlp::TY), 0, (cudaStream_t)stream>>>(
        x, batch, features, mean, var, running_mean, running_var, momentum);
    return launched("bn_stats");
}

extern "C" int gsr_bn_act_fwd_f32(const float* x, const float* mean, const float* var, const float* gamma,
                                  const float* beta, float eps, float p_drop, unsigned long long seed, int relu,
                                  float* y, int batch, int features, void* stream) {
    if (!x || !mean || !var || !gamma || !beta || !y || batch <= 0 || features <= 0 || !(p_drop >= 0.f && p_drop < 1.f))
        return fail(GSR_E_ARG, "bn_act_fwd: bad argument");
    const long long total = (long long)batch * features;
    mlp::bn_act_fwd_kernel<<<mlp::grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, mean, var, gamma, beta, eps, p_drop, seed,
                                                                                  relu, y, total, features);
    return launched("bn_act_fwd");
}

extern "C" int gsr_bn_act_bwd_f32(const float* x, const float* mean, const float* var, const float* gamma,
                                  const float* beta, const float* dy, float eps, float p_drop,
                                  unsigned long long seed, int relu, int batch_stats, float* dx, float* dgamma,
                                  float* dbeta, int batch, int features, void* stream) {
    if (!x || !mean || !var || !gamma || !beta || !dy || !dx || !dgamma || !dbeta || batch <= 0 || features <= 0 ||
        !(p_drop >= 0.f && p_drop < 1.f))
        return fail(GSR_E_ARG, "bn_act_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    mlp::bn_bwd_reduce_kernel<<<(features + mlp::TX - 1) / mlp::TX, dim3(mlp::TX, mlp::TY), 0, st>>>(
        x, mean, var, gamma, beta, dy, eps, p_drop, seed, relu, dgamma, dbeta, batch, features);
    int rc = launched("bn_bwd_reduce");
    if (rc) return rc;
    const long long total = (long long)batch * features;
    mlp::bn_bwd_apply_kernel<<<mlp::grid_for(total), 256, 0, st>>>(x, mean, var, gamma, beta, dy, dgamma, dbeta, eps, p_drop, seed,
                                                                  relu, batch_stats, dx, total, batch, features);
    return launched("bn_bwd_apply");
}
