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
// ncu on B200: the hardware float <-> double and int -> double conversions (F2F.F64.F32, I2F.F64) execute on the XU
// pipe at ~3 per clock and SM and BOUND these passes (sm__inst_executed_pipe_xu 120 % of its sustained peak, DRAM 38 %,
// FP64 pipe 18 % in pass 1).  Widening an fp32 to an fp64 is exact, so it is done with three integer operations
// instead; zero stays zero, denormals / Inf / NaN (exponent field 0 or 255: never in connectivity data) take the
// hardware conversion.
__device__ __forceinline__ double f2d(float x) {
    const uint32_t u = __float_as_uint(x);
    const uint32_t ex = u & 0x7F800000u;
    if (ex != 0u && ex != 0x7F800000u)
        return __hiloint2double((int)((u & 0x80000000u) | (((u & 0x7FFFFFFFu) >> 3) + 0x38000000u)), (int)(u << 29));
    if ((u & 0x7FFFFFFFu) == 0u) return __hiloint2double((int)u, 0);          // +-0
    return (double)x;
}
// (double)e for |e| < 2^31 with one DADD: 2^52 + 2^31 + e is exactly representable, its low word is e ^ 0x80000000
__device__ __forceinline__ double i2d(int e) {
    return __hiloint2double(0x43300000, (int)((uint32_t)e ^ 0x80000000u)) - 4503601774854144.0;
}
// 1 / x to full double precision for positive normal x: MUFU.RCP64H seed (>= 20 bits) and two Newton steps
__device__ __forceinline__ double rcp_pos(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-x, y, 2.0);
    y = y * fma(-x, y, 2.0);
    return y;
}

constexpr int kEvalBlocks = 148 * 8;
constexpr int kEvalSums = 8;

// Element visitor: strict upper triangle of [batch, n, n] (upper) or flat [batch, n].
// For the upper triangle a warp takes one matrix row at a time so that reads stay
// coalesced and only the needed half of every matrix is touched.
template <bool UPPER, typename F>
__device__ __forceinline__ void eval_visit(const float* __restrict__ p, const float* __restrict__ g,
                                           long long batch, int n, F f) {
    if (UPPER) {
        // A warp takes FOUR consecutive matrix rows at a time and walks their columns 32 at a time: the 8 loads of a
        // column block (4 rows x pred / gt) are all issued before the first value is used.  Only the strict upper
        // triangle is loaded: lanes left of the diagonal are predicated off.
        constexpr int R = 4;
        const long long rows = batch * n;
        const int lane = threadIdx.x & 31;
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long r0 = warp * R; r0 < rows; r0 += nwarps * R) {
            const int i0 = (int)(r0 % n);
            int first = i0 + 1;                                     // smallest column any of the rows needs ...
            if (i0 + R > n) first = 1;                              // ... (a group that wraps into the next matrix starts at row 0 again)
            for (int j0 = (first / 32) * 32; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                float pv[R], gv[R];
                bool ok[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int i = i0 + r;
                    if (i >= n) i -= n;
                    ok[r] = (r0 + r < rows) && j > i && j < n;
                    const long long off = (r0 + r) * n + j;
                    pv[r] = ok[r] ? __ldg(p + off) : 0.f;
                    gv[r] = ok[r] ? __ldg(g + off) : 0.f;
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (ok[r]) f(pv[r], gv[r]);
            }
        }
    } else {
        const long long total = batch * n;
        const long long stride = (long long)gridDim.x * blockDim.x;
        long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; e + 3 * stride < total; e += 4 * stride) {           // four independent loads in flight per operand
            const float p0 = __ldg(p + e), p1 = __ldg(p + e + stride), p2 = __ldg(p + e + 2 * stride), p3 = __ldg(p + e + 3 * stride);
            const float g0 = __ldg(g + e), g1 = __ldg(g + e + stride), g2 = __ldg(g + e + 2 * stride), g3 = __ldg(g + e + 3 * stride);
            f(p0, g0); f(p1, g1); f(p2, g2); f(p3, g3);
        }
        for (; e < total; e += stride) f(p[e], g[e]);
    }
}

template <bool UPPER>
__global__ void __launch_bounds__(kRedThreads)
eval_pass1_kernel(const float* __restrict__ p, const float* __restrict__ g, long long batch, int n,
                  double* __restrict__ partials /* [6][gridDim.x] */) {
    __shared__ double s_buf[kRedThreads / 32];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    eval_visit<UPPER>(p, g, batch, n, [&](float pf, float gf) {
        const double pv = f2d(pf), gv = f2d(gf);
        s0 += fabs(pv - gv); s1 += pv; s2 += gv; s3 += pv * pv; s4 += gv * gv; s5 += pv * gv;
    });
    double v[6] = {s0, s1, s2, s3, s4, s5};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double r = block_sum(v[k], s_buf);
        if (threadIdx.x == 0) partials[(size_t)k * gridDim.x + blockIdx.x] = r;
    }
}

// Double-precision logarithm for the Jensen-Shannon pass, ~12 FP64 operations instead of the ~40 instructions of the
// library's log(): x = 2^e m, m in [1, 2); the top 7 mantissa bits pick c_i = 1 + (i + 1/2) / 128 from a 128-entry
// table of (1 / c_i, log c_i); r = m / c_i - 1, |r| <= 2^-8, and log1p(r) is its degree-8 Taylor polynomial
// (truncation r^9 / 9 < 2^-75).  Absolute error ~2e-16 (the table entries and e ln 2 are correctly rounded doubles),
// for positive normal x -- the pass only calls it on ratios of positive fp32 data.
constexpr int kLogTab = 128;
struct LogTab { double2 e[kLogTab]; };          // (1 / c_i, log c_i): one 16-byte shared-memory load per logarithm

__global__ void eval_log_table_kernel(LogTab* __restrict__ t) {
    const int i = threadIdx.x;
    if (i < kLogTab) {
        const double c = 1.0 + (i + 0.5) / kLogTab;
        t->e[i] = make_double2(1.0 / c, log(c));
    }
}

__device__ __forceinline__ double log_tab(double x, const double2* __restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int e = (int)(bits >> 52) - 1023;
    const int i = (int)(bits >> 45) & (kLogTab - 1);
    const double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    const double2 t = tab[i];
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, 1.0 / 9.0, -1.0 / 8.0);
    q = fma(q, r, 1.0 / 7.0); q = fma(q, r, -1.0 / 6.0); q = fma(q, r, 1.0 / 5.0); q = fma(q, r, -1.0 / 4.0);
    q = fma(q, r, 1.0 / 3.0); q = fma(q, r, -0.5);
    q = fma(q * r, r, r);                                            // r - r^2/2 + r^3/3 - ...
    return fma(i2d(e), 0.693147180559945309417232121458, t.y + q);
}

// rel_entr terms of the rare elements with a zero (or non-positive) entry: scipy's conventions, out of line so that the
// streaming loop stays small enough for the instruction cache
__device__ __noinline__ void eval_rel_entr_slow(double a, double b, double& ka, double& kb) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double m = (a + b) / 2.0;
    if (a > 0 && m > 0) ka += a * log(a / m); else if (!(a == 0 && m >= 0)) ka += inf;
    if (b > 0 && m > 0) kb += b * log(b / m); else if (!(b == 0 && m >= 0)) kb += inf;
}

template <bool UPPER>
__global__ void __launch_bounds__(kRedThreads)
eval_pass2_kernel(const float* __restrict__ p, const float* __restrict__ g, long long batch, int n,
                  const double* __restrict__ sums, const LogTab* __restrict__ tab, double* __restrict__ partials /* [2][gridDim.x] */) {
    __shared__ double s_buf[kRedThreads / 32];
    __shared__ double2 s_tab[kLogTab];
    for (int i = threadIdx.x; i < kLogTab; i += blockDim.x) s_tab[i] = tab->e[i];
    __syncthreads();
    const double sp = sums[1], sg = sums[2];
    const double inv_sp = 1.0 / sp, inv_sg = 1.0 / sg;
    double ka = 0, kb = 0;
    // scipy.spatial.distance.jensenshannon(gt, pred): a = gt/sum, b = pred/sum, m = (a+b)/2,
    // rel_entr(x, m) = x log(x/m) for x > 0, m > 0; 0 for x == 0; +inf otherwise.
    eval_visit<UPPER>(p, g, batch, n, [&](float pf, float gf) {
        // a / m = 2 a / (a + b) and b / m = 2 b / (a + b) share ONE reciprocal of a + b (MUFU seed and two Newton
        // steps instead of a double division), and the two logarithms are the table-driven log_tab above.
        const double a = f2d(gf) * inv_sg, b = f2d(pf) * inv_sp;
        const double s2 = a + b;                                         // 2 m
        if (a > 0 && b > 0 && s2 > 1e-290 && s2 < 1e290) {               // the common case: both terms finite
            const double y = rcp_pos(s2);
            const double xa = 2.0 * a * y, xb = 2.0 * b * y;             // a / m, b / m in (0, 2)
            ka = fma(a, log_tab(xa, s_tab), ka);
            kb = fma(b, log_tab(xb, s_tab), kb);
        } else eval_rel_entr_slow(a, b, ka, kb);
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

extern "C" size_t gsr_eval_workspace(void) { return (size_t)kEvalSums * kEvalBlocks + kEvalSums + 2 * kLogTab; }

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
    LogTab* tab = reinterpret_cast<LogTab*>(workspace + (size_t)kEvalSums * kEvalBlocks + kEvalSums);
    eval_log_table_kernel<<<1, kLogTab, 0, st>>>(tab);
    if ((rc = launched("eval_log_table"))) return rc;
    if (upper) eval_pass2_kernel<true><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, sums, tab, partials);
    else       eval_pass2_kernel<false><<<kEvalBlocks, kRedThreads, 0, st>>>(pred, gt, batch, n, sums, tab, partials);
    if ((rc = launched("eval_pass2"))) return rc;
    eval_fold_kernel<<<1, kRedThreads, 0, st>>>(partials, sums + 6, 2, kEvalBlocks);
    if ((rc = launched("eval_fold2"))) return rc;
    const double count = upper ? (double)batch * ((double)n * (n - 1) / 2.0) : (double)batch * n;
    eval_finish_kernel<<<1, 1, 0, st>>>(sums, count, results_out);
    return launched("eval_finish");
}
