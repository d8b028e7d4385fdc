// SURVEY 8 f, row N2: the MLP baseline (models/mlp.py:7-27) between its Linear layers.
//   [Linear] -> BatchNorm1d -> Dropout(0.1) -> ReLU           (x is [batch][features], row-major)
// The Linear layers are dense GEMMs on gsr_tcgemm_f32 / gsr_bgemm_f32; the three layers in between are one
// memory-bound pass here (batch statistics first when training).  The dropout mask is a counter-based hash of
// (seed, element index): nothing is stored, the backward pass recomputes it, and the CPU restatement used by the
// tests reproduces it bit for bit.
#include <algorithm>

#include "common.cuh"

namespace gsr {
namespace mlp {

constexpr int TX = 32, TY = 32;        // column reductions: 32 features x 32 batch phases per CTA
constexpr int EW_THREADS = 128, EW_ROWS = 16;   // elementwise passes: a thread owns VEC columns over EW_ROWS rows

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

// A thread owns VEC consecutive features (float4 when the row length allows) and walks EW_ROWS batch rows: the
// per-feature constants are loaded and inverted once, the element index of the dropout hash needs no division.
template <int VEC> struct Pack { float v[VEC]; };
template <int VEC> __device__ __forceinline__ Pack<VEC> load_pack(const float* p) {
    Pack<VEC> r;
    if (VEC == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else r.v[0] = __ldg(p);
    return r;
}
template <int VEC> __device__ __forceinline__ void store_pack(float* p, const Pack<VEC>& r) {
    if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else p[0] = r.v[0];
}

template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ var, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, float p_drop,
                                                                unsigned long long seed, float slope, float* __restrict__ y,
                                                                int batch, int features) {
    const int h = (blockIdx.x * EW_THREADS + threadIdx.x) * VEC;
    if (h >= features) return;
    const Pack<VEC> m = load_pack<VEC>(mean + h), va = load_pack<VEC>(var + h), ga = load_pack<VEC>(gamma + h), be = load_pack<VEC>(beta + h);
    float rstd[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) rstd[i] = 1.f / sqrtf(va.v[i] + eps);
    const int b0 = blockIdx.y * EW_ROWS, b1 = min(b0 + EW_ROWS, batch);
    for (int b = b0; b < b1; ++b) {
        const size_t e = (size_t)b * features + h;
        Pack<VEC> v = load_pack<VEC>(x + e);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float z = (v.v[i] - m.v[i]) * rstd[i] * ga.v[i] + be.v[i];
            z *= keep_scale(p_drop, seed, (unsigned long long)(e + i));
            v.v[i] = z > 0.f ? z : slope * z;              // slope of the negative side: 0 = ReLU, 0.01 = LeakyReLU, 1 = none
        }
        store_pack<VEC>(y + e, v);
    }
}

// g = dy through ReLU and dropout; d beta = sum_b g, d gamma = sum_b g * xhat
__device__ __forceinline__ float upstream(float dy, float xhat, float gamma, float beta, float p_drop,
                                          unsigned long long seed, unsigned long long index, float slope) {
    const float k = keep_scale(p_drop, seed, index);
    return ((xhat * gamma + beta) * k > 0.f) ? dy * k : dy * k * slope;
}

__global__ void __launch_bounds__(TX * TY) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ var, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ dy,
                                                                float eps, float p_drop, unsigned long long seed, float relu,
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

template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                                  const float* __restrict__ var, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ dy,
                                                                  const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                                  float eps, float p_drop, unsigned long long seed, float relu,
                                                                  int batch_stats, float* __restrict__ dx, int batch, int features) {
    const int h = (blockIdx.x * EW_THREADS + threadIdx.x) * VEC;
    if (h >= features) return;
    const Pack<VEC> m = load_pack<VEC>(mean + h), va = load_pack<VEC>(var + h), ga = load_pack<VEC>(gamma + h), be = load_pack<VEC>(beta + h);
    const Pack<VEC> dg = load_pack<VEC>(dgamma + h), db = load_pack<VEC>(dbeta + h);
    const float inv_b = 1.f / (float)batch;
    float rstd[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) rstd[i] = 1.f / sqrtf(va.v[i] + eps);
    const int b0 = blockIdx.y * EW_ROWS, b1 = min(b0 + EW_ROWS, batch);
    for (int b = b0; b < b1; ++b) {
        const size_t e = (size_t)b * features + h;
        const Pack<VEC> xv = load_pack<VEC>(x + e), dv = load_pack<VEC>(dy + e);
        Pack<VEC> out;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float xhat = (xv.v[i] - m.v[i]) * rstd[i];
            float g = upstream(dv.v[i], xhat, ga.v[i], be.v[i], p_drop, seed, (unsigned long long)(e + i), relu);
            if (batch_stats) g -= (db.v[i] + xhat * dg.v[i]) * inv_b;                     // the statistics depend on x
            out.v[i] = ga.v[i] * rstd[i] * g;
        }
        store_pack<VEC>(dx + e, out);
    }
}

// training_template.ipynb:176-180 (notebook MLP): x = leaky_relu(block(x) + residual) -- out = act(a + b), and its
// backward (the same gradient flows to both addends); :132 nn.Sigmoid on the output layer and its backward.
__global__ void __launch_bounds__(256) add_act_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                          long long count, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) {
        const float z = __ldg(a + e) + __ldg(b + e);
        out[e] = z > 0.f ? z : slope * z;
    }
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ out, const float* __restrict__ g, float* __restrict__ gin,
                                                      long long count, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
        gin[e] = __ldg(out + e) > 0.f ? __ldg(g + e) : slope * __ldg(g + e);      // slope > 0: sign(out) = sign(a + b)
}
__global__ void __launch_bounds__(256) sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) y[e] = 1.0f / (1.0f + expf(-__ldg(x + e)));
}
__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ gx, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) {
        const float v = __ldg(y + e);
        gx[e] = __ldg(g + e) * v * (1.0f - v);
    }
}

// ---- spectral_norm(nn.Linear) (training_template.ipynb:125,134,143; torch.nn.utils.spectral_norm.compute_weight):
//   W_eff = W / sigma with sigma = u . (W v);   autograd:  dW = G / sigma + dsigma u v^T,  dsigma = -sum(G o W_eff) / sigma.
// torch's hook runs these as ~6 broadcast elementwise / reduction launches per layer and direction (the non-vectorised
// TensorIterator path: 0.6 ms of the MLP step at 51.7 M weights); here: one streaming pass forward, a deterministic
// two-stage dot and one pass backward.  sigma stays a DEVICE scalar (no host synchronisation).
constexpr int kSpecBlocks = 148 * 8;
__global__ void __launch_bounds__(256) scale_div_kernel(const float* __restrict__ x, const float* __restrict__ sigma, float* __restrict__ y, long long count) {
    const float s = __ldg(sigma);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) y[e] = __ldg(x + e) / s;
}
__global__ void __launch_bounds__(256) scale_div4_kernel(const float4* __restrict__ x, const float* __restrict__ sigma, float4* __restrict__ y, long long count4) {
    const float s = __ldg(sigma);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count4; e += stride) {
        const float4 v = ldg_stream(x + e);
        y[e] = make_float4(v.x / s, v.y / s, v.z / s, v.w / s);
    }
}
// partial[block] = sum over the block's elements of g * w (fp32 products, fp64 sums, fixed order)
__global__ void __launch_bounds__(256) spectral_dot_kernel(const float* __restrict__ g, const float* __restrict__ w, long long count, double* __restrict__ partial) {
    __shared__ double s_w[8];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) acc += (double)(__ldg(g + e) * __ldg(w + e));
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) spectral_fold_kernel(double* __restrict__ partial, int nblocks) {      // total -> partial[nblocks]
    __shared__ double s_w[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc += partial[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        partial[nblocks] = t;
    }
}
// dW[i][j] = G[i][j] / sigma + dsigma u[i] v[j]; one CTA per row
__global__ void __launch_bounds__(256) spectral_apply_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                                                             const float* __restrict__ sigma, const double* __restrict__ total,
                                                             float* __restrict__ dw, int cols, int vec) {
    const float s = __ldg(sigma);
    const float du = (float)(-(*total) / (double)s) * __ldg(u + blockIdx.x);
    const float* gr = g + (size_t)blockIdx.x * cols;
    float* dr = dw + (size_t)blockIdx.x * cols;
    if (vec) {
        for (int j = 4 * threadIdx.x; j < cols; j += 4 * 256) {
            const float4 a = ldg_stream(reinterpret_cast<const float4*>(gr + j)), b = __ldg(reinterpret_cast<const float4*>(v + j));
            *reinterpret_cast<float4*>(dr + j) = make_float4(fmaf(du, b.x, a.x / s), fmaf(du, b.y, a.y / s), fmaf(du, b.z, a.z / s), fmaf(du, b.w, a.w / s));
        }
    } else {
        for (int j = threadIdx.x; j < cols; j += 256) dr[j] = fmaf(du, __ldg(v + j), __ldg(gr + j) / s);
    }
}
static int ew_blocks(long long count) { return (int)std::min<long long>((count + 255) / 256, 148LL * 16); }

static bool vec4_ok(int features, const void* a, const void* b, const void* c) {
    return features % 4 == 0 && aligned16(a) && aligned16(b) && (!c || aligned16(c));
}

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
                                  const float* beta, float eps, float p_drop, unsigned long long seed, float relu,
                                  float* y, int batch, int features, void* stream) {
    if (!x || !mean || !var || !gamma || !beta || !y || batch <= 0 || features <= 0 || !(p_drop >= 0.f && p_drop < 1.f))
        return fail(GSR_E_ARG, "bn_act_fwd: bad argument");
    const bool v4 = mlp::vec4_ok(features, x, y, nullptr) && aligned16(mean) && aligned16(var) && aligned16(gamma) && aligned16(beta);
    const int per = mlp::EW_THREADS * (v4 ? 4 : 1);
    dim3 grid((features + per - 1) / per, (batch + mlp::EW_ROWS - 1) / mlp::EW_ROWS);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "bn_act_fwd: batch too large");
    if (v4) mlp::bn_act_fwd_kernel<4><<<grid, mlp::EW_THREADS, 0, (cudaStream_t)stream>>>(x, mean, var, gamma, beta, eps, p_drop, seed, relu, y, batch, features);
    else mlp::bn_act_fwd_kernel<1><<<grid, mlp::EW_THREADS, 0, (cudaStream_t)stream>>>(x, mean, var, gamma, beta, eps, p_drop, seed, relu, y, batch, features);
    return launched("bn_act_fwd");
}

extern "C" int gsr_bn_act_bwd_f32(const float* x, const float* mean, const float* var, const float* gamma,
                                  const float* beta, const float* dy, float eps, float p_drop,
                                  unsigned long long seed, float relu, int batch_stats, float* dx, float* dgamma,
                                  float* dbeta, int batch, int features, void* stream) {
    if (!x || !mean || !var || !gamma || !beta || !dy || !dx || !dgamma || !dbeta || batch <= 0 || features <= 0 ||
        !(p_drop >= 0.f && p_drop < 1.f))
        return fail(GSR_E_ARG, "bn_act_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    mlp::bn_bwd_reduce_kernel<<<(features + mlp::TX - 1) / mlp::TX, dim3(mlp::TX, mlp::TY), 0, st>>>(
        x, mean, var, gamma, beta, dy, eps, p_drop, seed, relu, dgamma, dbeta, batch, features);
    int rc = launched("bn_bwd_reduce");
    if (rc) return rc;
    const bool v4 = mlp::vec4_ok(features, x, dy, dx) && aligned16(mean) && aligned16(var) && aligned16(gamma) && aligned16(beta) &&
                    aligned16(dgamma) && aligned16(dbeta);
    const int per = mlp::EW_THREADS * (v4 ? 4 : 1);
    dim3 grid((features + per - 1) / per, (batch + mlp::EW_ROWS - 1) / mlp::EW_ROWS);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "bn_act_bwd: batch too large");
    if (v4) mlp::bn_bwd_apply_kernel<4><<<grid, mlp::EW_THREADS, 0, st>>>(x, mean, var, gamma, beta, dy, dgamma, dbeta, eps, p_drop, seed, relu, batch_stats, dx, batch, features);
    else mlp::bn_bwd_apply_kernel<1><<<grid, mlp::EW_THREADS, 0, st>>>(x, mean, var, gamma, beta, dy, dgamma, dbeta, eps, p_drop, seed, relu, batch_stats, dx, batch, features);
    return launched("bn_bwd_apply");
}

extern "C" int gsr_add_act_fwd_f32(const float* a, const float* b, float* out, long long count, float negative_slope, void* stream) {
    if (!a || !b || !out || count < 0) return fail(GSR_E_ARG, "add_act_fwd: bad argument");
    if (count == 0) return 0;
    mlp::add_act_fwd_kernel<<<mlp::ew_blocks(count), 256, 0, (cudaStream_t)stream>>>(a, b, out, count, negative_slope);
    return launched("add_act_fwd");
}
extern "C" int gsr_act_bwd_f32(const float* out, const float* grad_out, float* grad_in, long long count, float negative_slope, void* stream) {
    if (!out || !grad_out || !grad_in || count < 0) return fail(GSR_E_ARG, "act_bwd: bad argument");
    if (count == 0) return 0;
    mlp::act_bwd_kernel<<<mlp::ew_blocks(count), 256, 0, (cudaStream_t)stream>>>(out, grad_out, grad_in, count, negative_slope);
    return launched("act_bwd");
}
extern "C" int gsr_sigmoid_fwd_f32(const float* x, float* y, long long count, void* stream) {
    if (!x || !y || count < 0) return fail(GSR_E_ARG, "sigmoid_fwd: bad argument");
    if (count == 0) return 0;
    mlp::sigmoid_fwd_kernel<<<mlp::ew_blocks(count), 256, 0, (cudaStream_t)stream>>>(x, y, count);
    return launched("sigmoid_fwd");
}
extern "C" int gsr_sigmoid_bwd_f32(const float* y, const float* grad_y, float* grad_x, long long count, void* stream) {
    if (!y || !grad_y || !grad_x || count < 0) return fail(GSR_E_ARG, "sigmoid_bwd: bad argument");
    if (count == 0) return 0;
    mlp::sigmoid_bwd_kernel<<<mlp::ew_blocks(count), 256, 0, (cudaStream_t)stream>>>(y, grad_y, grad_x, count);
    return launched("sigmoid_bwd");
}

extern "C" int gsr_scale_div_f32(const float* x, const float* sigma, float* y, long long count, void* stream) {
    if (!x || !sigma || !y || count < 0) return fail(GSR_E_ARG, "scale_div: bad argument");
    if (count == 0) return 0;
    if (count % 4 == 0 && aligned16(x) && aligned16(y))
        mlp::scale_div4_kernel<<<mlp::ew_blocks(count / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), sigma, reinterpret_cast<float4*>(y), count / 4);
    else mlp::scale_div_kernel<<<mlp::ew_blocks(count), 256, 0, (cudaStream_t)stream>>>(x, sigma, y, count);
    return launched("scale_div");
}
extern "C" size_t gsr_spectral_bwd_workspace(void) { return (size_t)mlp::kSpecBlocks + 1; }      /* doubles */
extern "C" int gsr_spectral_bwd_f32(const float* grad_w_eff, const float* w_eff, const float* u, const float* v, const float* sigma,
                                    float* grad_w, int rows, int cols, double* workspace, void* stream) {
    if (!grad_w_eff || !w_eff || !u || !v || !sigma || !grad_w || !workspace || rows <= 0 || cols <= 0)
        return fail(GSR_E_ARG, "spectral_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long count = (long long)rows * cols;
    const int blocks = (int)std::min<long long>((count + 255) / 256, mlp::kSpecBlocks);
    mlp::spectral_dot_kernel<<<blocks, 256, 0, st>>>(grad_w_eff, w_eff, count, workspace);
    int rc = launched("spectral_dot");
    if (rc) return rc;
    mlp::spectral_fold_kernel<<<1, 256, 0, st>>>(workspace, blocks);
    if ((rc = launched("spectral_fold"))) return rc;
    const int vec = cols % 4 == 0 && aligned16(grad_w_eff) && aligned16(grad_w) && aligned16(v);
    mlp::spectral_apply_kernel<<<rows, 256, 0, st>>>(grad_w_eff, u, v, sigma, workspace + blocks, grad_w, cols, vec);
    return launched("spectral_apply");
}
