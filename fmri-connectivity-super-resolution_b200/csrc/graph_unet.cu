// Graph U-Net pooling ops (models/ops.py): score, top-k select, gather * score, unpool
// scatter and their autograd.  All of them are subject independent on the GSR-Net path
// (the U-Net is fed the identity), tiny (n <= 160 rows of 268 floats) and
// latency bound: one launch each, no host round trips, indices stay on the device as
// int64 (torch's index dtype) so they can be handed straight back to the caller.
#include "common.cuh"

namespace gsr {

// models/ops.py:35-38: one warp per row
__global__ void pool_scores_kernel(const float* __restrict__ X, const float* __restrict__ p,
                                   const float* __restrict__ c, float* __restrict__ scores, int n, int d) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) acc = fmaf(__ldg(X + (size_t)row * d + k), __ldg(p + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        const float s = (acc + __ldg(c)) / 100.0f;
        scores[row] = 1.0f / (1.0f + expf(-s));
    }
}

// torch.topk order: NaN first, then descending value; ties by ascending index (stated rule)
__device__ __forceinline__ bool topk_before(float sj, int j, float si, int i) {
    const bool nj = isnan(sj), ni = isnan(si);
    if (nj || ni) return nj && (!ni || j < i);
    return sj > si || (sj == si && j < i);
}

__global__ void pool_topk_kernel(const float* __restrict__ scores, int n, int k,
                                 int64_t* __restrict__ idx, float* __restrict__ values) {
    extern __shared__ float s[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = scores[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float si = s[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += topk_before(s[j], j, si, i) ? 1 : 0;
        if (rank < k) { idx[rank] = i; values[rank] = si; }
    }
}

// models/ops.py:43-45
__global__ void pool_gather_scale_kernel(const float* __restrict__ X, const int64_t* __restrict__ idx,
                                         const float* __restrict__ values, float* __restrict__ out, int d) {
    const int t = blockIdx.x;
    const size_t src = (size_t)idx[t] * d;
    const float v = values[t];
    for (int c = threadIdx.x; c < d; c += blockDim.x) out[(size_t)t * d + c] = X[src + c] * v;
}

// new_X[t] = X[i] * sigma_i with i = idx[t], sigma = sigmoid((X.p + c)/100):
//   dX[i] = g[t] * sigma_i + ds_i * p,  ds_i = (g[t] . X[i]) * sigma_i (1 - sigma_i) / 100
__global__ void pool_bwd_rows_kernel(const float* __restrict__ X, const float* __restrict__ p,
                                     const float* __restrict__ scores, const int64_t* __restrict__ idx,
                                     const float* __restrict__ g, float* __restrict__ dX,
                                     float* __restrict__ ds, int d, int n) {
    __shared__ float s_part[32];
    __shared__ float s_ds;
    const int t = blockIdx.x;
    const int64_t i64 = idx[t];
    if (i64 < 0 || i64 >= n) return;                       // uniform over the block
    const int i = (int)i64;
    const float sig = scores[i];
    float acc = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) acc = fmaf(g[(size_t)t * d + c], X[(size_t)i * d + c], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_part[w];
        const float v = tot * (sig * (1.0f - sig)) / 100.0f;
        s_ds = v;
        ds[i] = v;
    }
    __syncthreads();
    const float dsi = s_ds;
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        dX[(size_t)i * d + c] = fmaf(dsi, p[c], g[(size_t)t * d + c] * sig);
}

// dp[c] = sum_i ds_i X[i][c];  dc = sum_i ds_i   (fixed order: 32 columns x 8 row phases per CTA, the phases folded in
// order -- one thread per column walking all n rows was a chain of n dependent FMAs, 9-19 us on the step's tail)
__global__ void __launch_bounds__(256) pool_bwd_proj_kernel(const float* __restrict__ X, const float* __restrict__ ds,
                                                            float* __restrict__ dp, float* __restrict__ dc, int n, int d) {
    __shared__ float part[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + x;
    float acc = 0.f;
    if (c < d)
        for (int i = y; i < n; i += 8) acc = fmaf(ds[i], X[(size_t)i * d + c], acc);
    part[y][x] = acc;
    __syncthreads();
    if (y == 0 && c < d) {
        float s = part[0][x];
#pragma unroll
        for (int i = 1; i < 8; ++i) s += part[i][x];
        dp[c] = s;
    }
    if (blockIdx.x == 0 && y == 1) {
        float a = 0.f;
        for (int i = x; i < n; i += 32) a += ds[i];
        a = warp_sum(a);
        if (x == 0) dc[0] = a;
    }
}

// models/ops.py:46-47
__global__ void gather_rows_cols_kernel(const float* __restrict__ A, const int64_t* __restrict__ idx,
                                        float* __restrict__ out, int n, int k) {
    const float* Ab = A + (size_t)blockIdx.y * n * n;
    float* ob = out + (size_t)blockIdx.y * k * k;
    const int r = blockIdx.x;
    const int64_t ir = idx[r];
    const bool row_ok = ir >= 0 && ir < n;                 // an index outside [0, n) never becomes an address (torch raises;
    const size_t src = (size_t)(row_ok ? ir : 0) * n;      // the Python wrappers validate caller-supplied indices)
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        const int64_t ic = idx[c];
        ob[(size_t)r * k + c] = (row_ok && ic >= 0 && ic < n) ? Ab[src + ic] : 0.f;
    }
}

// models/ops.py:22-23 scatter (rows named by idx are distinct) and its autograd (gather)
__global__ void scatter_rows_kernel(const float* __restrict__ X, const int64_t* __restrict__ idx,
                                    float* __restrict__ out, int d, int n) {
    const int64_t i = idx[blockIdx.x];
    if (i < 0 || i >= n) return;                           // out of range: no write (never an out-of-bounds store)
    const size_t dst = (size_t)i * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) out[dst + c] = X[(size_t)blockIdx.x * d + c];
}
__global__ void gather_rows_kernel(const float* __restrict__ X, const int64_t* __restrict__ idx,
                                   float* __restrict__ out, int d) {
    const size_t src = (size_t)idx[blockIdx.x] * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) out[(size_t)blockIdx.x * d + c] = X[src + c];
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_pool_scores(const float* X, const float* p, const float* c, float* scores,
                               int n, int d, void* stream) {
    if (!X || !p || !c || !scores || n <= 0 || d <= 0) return fail(GSR_E_ARG, "pool_scores: bad argument");
    pool_scores_kernel<<<(n * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(X, p, c, scores, n, d);
    return launched("pool_scores");
}

extern "C" int gsr_pool_topk(const float* scores, int n, int k, int64_t* idx, float* values, void* stream) {
    if (!scores || !idx || !values || n <= 0 || k <= 0 || k > n) return fail(GSR_E_ARG, "pool_topk: bad argument");
    if (n > 4096) return fail(GSR_E_UNSUPPORTED, "pool_topk: n > 4096");
    const int threads = n < 256 ? ((n + 31) & ~31) : 256;
    pool_topk_kernel<<<1, threads, n * sizeof(float), (cudaStream_t)stream>>>(scores, n, k, idx, values);
    return launched("pool_topk");
}

extern "C" int gsr_pool_gather_scale_fwd(const float* X, const int64_t* idx, const float* values,
                                         float* new_X, int k, int d, void* stream) {
    if (!X || !idx || !values || !new_X || k <= 0 || d <= 0) return fail(GSR_E_ARG, "pool_gather_scale_fwd: bad argument");
    pool_gather_scale_kernel<<<k, 128, 0, (cudaStream_t)stream>>>(X, idx, values, new_X, d);
    return launched("pool_gather_scale");
}

extern "C" int gsr_pool_bwd(const float* X, const float* p, const float* scores, const int64_t* idx,
                            const float* grad_new_X, float* grad_X, float* grad_p, float* grad_c,
                            float* ds_workspace, int n, int k, int d, void* stream) {
    if (!X || !p || !scores || !idx || !grad_new_X || !grad_X || !grad_p || !grad_c || !ds_workspace ||
        n <= 0 || k <= 0 || k > n || d <= 0) return fail(GSR_E_ARG, "pool_bwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_X, 0, (size_t)n * d * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ds_workspace, 0, (size_t)n * sizeof(float), st);
    if (e != cudaSuccess) return fail((int)e, "pool_bwd: memset failed");
    pool_bwd_rows_kernel<<<k, 128, 0, st>>>(X, p, scores, idx, grad_new_X, grad_X, ds_workspace, d, n);
    int rc = launched("pool_bwd_rows");
    if (rc) return rc;
    pool_bwd_proj_kernel<<<(d + 31) / 32, 256, 0, st>>>(X, ds_workspace, grad_p, grad_c, n, d);
    return launched("pool_bwd_proj");
}

extern "C" int gsr_gather_rows_cols(const float* A, const int64_t* idx, float* out,
                                    int batch, int n, int k, void* stream) {
    if (!A || !idx || !out || batch < 0 || n <= 0 || k <= 0) return fail(GSR_E_ARG, "gather_rows_cols: bad argument");
    if (batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "gather_rows_cols: batch > 65535");
    gather_rows_cols_kernel<<<dim3(k, batch), 128, 0, (cudaStream_t)stream>>>(A, idx, out, n, k);
    return launched("gather_rows_cols");
}

extern "C" int gsr_unpool_fwd(const float* X, const int64_t* idx, float* new_X, int n, int k, int d, void* stream) {
    if (!X || !idx || !new_X || n <= 0 || k <= 0 || d <= 0) return fail(GSR_E_ARG, "unpool_fwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(new_X, 0, (size_t)n * d * sizeof(float), st);
    if (e != cudaSuccess) return fail((int)e, "unpool_fwd: memset failed");
    scatter_rows_kernel<<<k, 128, 0, st>>>(X, idx, new_X, d, n);
    return launched("unpool_scatter");
}

extern "C" int gsr_unpool_bwd(const float* grad_new_X, const int64_t* idx, float* grad_X, int k, int d, void* stream) {
    if (!grad_new_X || !idx || !grad_X || k <= 0 || d <= 0) return fail(GSR_E_ARG, "unpool_bwd: bad argument");
    gather_rows_kernel<<<k, 128, 0, (cudaStream_t)stream>>>(grad_new_X, idx, grad_X, d);
    return launched("unpool_gather");
}
