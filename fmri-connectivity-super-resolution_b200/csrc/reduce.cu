// Loss / metric reductions of the GSR-Net path:
//   nn.L1Loss (utils/gsr_train.py:9,39-43) forward + backward, and the
//   MAE / Pearson / Jensen-Shannon sweep of utils/evaluation.py:96-112.
// All of them are single-pass streaming reads (HBM-bound), fp64 accumulation,
// deterministic two-stage reductions (per-CTA partials, then a fixed-order sum).
#include "common.cuh"

namespace gsr {

constexpr int kRedThreads = 256;
constexpr long long kL1Chunk = 16384;     // elements per CTA in the L1 kernels

__device__ __forceinline__ double block_sum(double v, double* s_buf) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < (blockDim.x >> 5)) ? s_buf[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;   // valid in warp 0
}

// grid = (chunks, batch): partial[b][chunk] = sum |x - y| over the chunk
__global__ void __launch_bounds__(kRedThreads)
l1_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long long y_stride,
                  double* __restrict__ partials, long long count) {
    __shared__ double s_buf[kRedThreads / 32];
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * count;
    const float* yb = y + (size_t)b * y_stride;
    const long long lo = (long long)blockIdx.x * kL1Chunk;
    const long long hi = min(lo + kL1Chunk, count);
    float acc = 0.f;       // <= 64 terms per thread: fp32 partial, fp64 across threads
    for (long long e = lo + threadIdx.x; e < hi; e += kRedThreads) acc += fabsf(xb[e] - yb[e]);
    const double r = block_sum((double)acc, s_buf);
    if (threadIdx.x == 0) partials[(size_t)b * gridDim.x + blockIdx.x] = r;
}

__global__ void l1_final_kernel(const double* __restrict__ partials, float* __restrict__ loss,
                                int chunks, long long count) {
    __shared__ double s_buf[kRedThreads / 32];
    const int b = blockIdx.x;
    double acc = 0.0;
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) acc += partials[(size_t)b * chunks + c];
    const double r = block_sum(acc, s_buf);
    if (threadIdx.x == 0) loss[b] = (float)(r / (double)count);
}

__global__ void __launch_bounds__(kRedThreads)
l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long y_stride,
              const float* __restrict__ grad_loss, float* __restrict__ gx, long long count) {
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * count;
    const float* yb = y + (size_t)b * y_stride;
    float* gb = gx + (size_t)b * count;
    const float scale = grad_loss[b] / (float)count;
    const long long lo = (long long)blockIdx.x * kL1Chunk;
    const long long hi = min(lo + kL1Chunk, count);
    for (long long e = lo + threadIdx.x; e < hi; e += kRedThreads) gb[e] = sgnf(xb[e] - yb[e]) * scale;
}

// y shared by the subjects (|W - U_hr|, utils/gsr_train.py:41): grad_y[i] = - sum_b sgn(x[b,i] - y[i]) grad_loss[b] / count
// in one pass over x -- threads (tx, ty) = (element, batch phase), fixed summation order -- instead of writing the
// [batch, count] gradient of x and summing it over the batch.
constexpr int kSharedTX = 32, kSharedTY = 32;
__global__ void __launch_bounds__(kSharedTX * kSharedTY)
l1_bwd_shared_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ grad_loss,
                     float* __restrict__ gy, int batch, long long count) {
    __shared__ float red[kSharedTY][kSharedTX + 1];
    const long long i = (long long)blockIdx.x * kSharedTX + threadIdx.x;
    const bool live = i < count;
    const float yi = live ? __ldg(y + i) : 0.f;
    float acc = 0.f;
    for (int b = threadIdx.y; b < batch; b += kSharedTY)
        if (live) acc += sgnf(__ldg(x + (size_t)b * count + i) - yi) * __ldg(grad_loss + b);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && live) {
        float total = 0.f;
#pragma unroll
        for (int j = 0; j < kSharedTY; ++j) total += red[j][threadIdx.x];
        gy[i] = -total / (float)count;
    }
}

// ---------------------------------------------------------------- evaluation sweep
constexpr int kEvalBlocks = 148 * 8;
constexpr int kEvalSums = 8;

// Element visitor: strict upper triangle of [batch, n, n] (upper) or flat [batch, n].
// For the upper triangle a warp takes one matrix row at a time so that reads stay
// coalesced and only the needed half of every matrix is touched.
template <bool UPPER, typename F>
__device__ __forceinline__ void eval_visit(const float* __restrict__ p, const float* __restrict__ g,
                                           long long batch, int n, F f) {
    if (UPPER) {
        const long long rows = batch * n;
        const int lane = threadIdx.x & 31;
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long r = warp; r < rows; r += nwarps) {
            const int i = (int)(r % n);
            const float* pr = p + r * n;
            const float* gr = g + r * n;
            for (int j = i + 1 + lane; j < n; j += 32) f(pr[j], gr[j]);
        }
    } else {
        const long long total = batch * n;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (long long)gridDim.x * blockDim.x) f(p[e], g[e]);
    }
}

template <bool UPPER>
__global__ void __launch_bounds__(kRedThreads)
eval_pass1_kernel(const float* __restrict__ p, const float* __restrict__ g, long long batch, int n,
                  double* __restrict__ partials /* [6][gridDim.x] */) {
    __shared__ double s_buf[kRedThreads / 32];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    eval_visit<UPPER>(p, g, batch, n, [&](float pf, float gf) {
        const double pv = pf, gv = gf;
        s0 += fabs(pv - gv); s1 += pv; s2 += gv; s3 += pv * pv; s4 += gv * gv; s5 += pv * gv;
    });
    double v[6] = {s0, s1, s2, s3, s4, s5};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double r = block_sum(v[k], s_buf);
        if (threadIdx.x == 0) partials[(size_t)k * gridDim.x + blockIdx.x] = r;
    }
}

template <bool UPPER>
__global__ void __launch_bounds__(kRedThreads)
eval_pass2_kernel(const float* __restrict__ p, const float* __restrict__ g, long long batch, int n,
                  const double* __restrict__ sums, double* __restrict__ partials /* [2][gridDim.x] */) {
    __shared__ double s_buf[kRedThreads / 32];
    const double sp = sums[1], sg = sums[2];
    const double inv_sp = 1.0 / sp, inv_sg = 1.0 / sg;
    double ka = 0, kb = 0;
    // scipy.spatial.distance.jensenshannon(gt, pred): a = gt/sum, b = pred/sum, m = (a+b)/2,
    // rel_entr(x, m) = x log(x/m) for x > 0, m > 0; 0 for x == 0; +inf otherwise.
    eval_visit<UPPER>(p, g, batch, n, [&](float pf, float gf) {
        // The kernel is bound by the FP64 pipe (two double logarithms per element), not by HBM: one division
        // instead of four -- a / m = 2 a / (a + b) and b / m = 2 b / (a + b) share the reciprocal of a + b.
        const double a = (double)gf * inv_sg, b = (double)pf * inv_sp;
        const double m = (a + b) / 2.0;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        const double inv_m = 1.0 / m;
        if (a > 0 && m > 0) ka += a * log(a * inv_m); else if (!(a == 0 && m >= 0)) ka += inf;
        if (b > 0 && m > 0) kb += b * log(b * inv_m); else if (!(b == 0 && m >= 0)) kb += inf;
    });
    double r = block_sum(ka, s_buf);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
    r = block_sum(kb, s_buf);
    if (threadIdx.x == 0) partials[gridDim.x + blockIdx.x] = r;
}

// sums[k] = sum over blocks of partials[k][*]  (one CTA, fixed order)
__global__ void eval_fold_kernel(const double* __restrict__ partials, double* __restrict__ sums,
                                 int nsums, int nblocks) {
    __shared__ double s_buf[kRedThreads / 32];
    for (int k = 0; k < nsums; ++k) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partials[(size_t)k * nblocks + i];
        const double r = block_sum(acc, s_buf);
        if (threadIdx.x == 0) sums[k] = r;
    }
}

__global__ void eval_finish_kernel(const double* __restrict__ sums, double count, double* __restrict__ out) {
    const double sabs = sums[0], sp = sums[1], sg = sums[2], spp = sums[3], sgg = sums[4], spg = sums[5];
    out[0] = sabs / count;
    const double cov = spg - sp * sg / count;
    const double vp = spp - sp * sp / count, vg = sgg - sg * sg / count;
    out[1] = cov / sqrt(vp * vg);
    out[2] = sqrt((sums[6] + sums[7]) / 2.0);
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_l1_partials(long long count) { return (int)((count + kL1Chunk - 1) / kL1Chunk); }

extern "C" int gsr_l1_loss_fwd(const float* x, const float* y, long long y_stride, float* loss_out,
                               double* partials, int batch, long long count, void* stream) {
    if (!x || !y || !loss_out || !partials || batch < 0 || count <= 0) return fail(GSR_E_ARG, "l1_loss_fwd: bad argument");
    if (batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "l1_loss_fwd: batch > 65535");
    const int chunks = gsr_l1_partials(count);
    l1_partial_kernel<<<dim3(chunks, batch), kRedThreads, 0, (cudaStream_t)stream>>>(x, y, y_stride, partials, count);
    int rc = launched("l1_partial");
    if (rc) return rc;
    l1_final_kernel<<<batch, kRedThreads, 0, (cudaStream_t)stream>>>(partials, loss_out, chunks, count);
    return launched("l1_final");
}

extern "C" int gsr_l1_loss_bwd(const float* x, const float* y, long long y_stride, const float* grad_loss,
                               float* grad_x, int batch, long long count, void* stream) {
    if (!x || !y || !grad_loss || !grad_x || batch < 0 || count <= 0) return fail(GSR_E_ARG, "l1_loss_bwd: bad argument");
    if (batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "l1_loss_bwd: batch > 65535");
    const int chunks = gsr_l1_partials(count);
    l1_bwd_kernel<<<dim3(chunks, batch), kRedThreads, 0, (cudaStream_t)stream>>>(x, y, y_stride, grad_loss, grad_x, count);
    return launched("l1_bwd");
}

extern "C" int gsr_l1_loss_bwd_shared_y(const float* x, const float* y, const float* grad_loss, float* grad_y,
                                       int batch, long long count, void* stream) {
    if (!x || !y || !grad_loss || !grad_y || batch < 0 || count <= 0) return fail(GSR_E_ARG, "l1_loss_bwd_shared_y: bad argument");
    const long long blocks = (count + kSharedTX - 1) / kSharedTX;
    if (blocks > 0x7fffffffLL) return fail(GSR_E_UNSUPPORTED, "l1_loss_bwd_shared_y: count too large");
    l1_bwd_shared_kernel<<<(unsigned)blocks, dim3(kSharedTX, kSharedTY), 0, (cudaStream_t)stream>>>(x, y, grad_loss, grad_y, batch, count);
    return launched("l1_bwd_shared_y");
}

extern "C" size_t gsr_eval_workspace(void) { return (size_t)kEvalSums * kEvalBlocks + kEvalSums; }

extern "C" int gsr_eval_reduce(const float* pred, const float* gt, long long batch, int n, int upper,
                               double* workspace, double* results_out, void* stream) {
    if (!pred || !gt || !workspace || !results_out || batch <= 0 || n <= 0) return fail(GSR_E_ARG, "eval_reduce: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = workspace;
    double* sums = workspace + (size_t)kEvalSums * kEvalBlocks;
    if (upper) eval_pass1_kernel<true><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, partials);
    else       eval_pass1_kernel<false><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, partials);
    int rc = launched("eval_pass1");
    if (rc) return rc;
    eval_fold_kernel<<<1, kRedThreads, 0, st>>>(partials, sums, 6, kEvalBlocks);
    if ((rc = launched("eval_fold1"))) return rc;
    if (upper) eval_pass2_kernel<true><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, sums, partials);
    else       eval_pass2_kernel<false><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, sums, partials);
    if ((rc = launched("eval_pass2"))) return rc;
    eval_fold_kernel<<<1, kRedThreads, 0, st>>>(partials, sums + 6, 2, kEvalBlocks);
    if ((rc = launched("eval_fold2"))) return rc;
    const double count = upper ? (double)batch * ((double)n * (n - 1) / 2.0) : (double)batch * n;
    eval_finish_kernel<<<1, 1, 0, st>>>(sums, count, results_out);
    return launched("eval_finish");
}
