// Memory-bound n x n kernels of the GSR-Net path: adjacency normalisation (fwd/bwd,
// optionally fused with abs + unit diagonal), symmetrise + unit diagonal + abs
// (fwd/bwd).  One CTA per subject matrix.  Every kernel walks the matrix in
// 32-row strips: the rows of the strip are read straight from global memory
// (coalesced), the matching 32-column strip (the transposed operand) is staged
// through padded shared memory so that both the global reads and the global
// writes stay coalesced.
#include "common.cuh"

namespace gsr {

constexpr int kEwThreads = 1024;   // one or two CTAs per SM: a matrix re-read by its CTA is still in L2
constexpr int kEwWarps = kEwThreads / 32;
constexpr int kStripPad = 33;   // 32 + 1: conflict-free transposed reads

// M(i,j) for pre = 1 is |mx(i,j)| with a unit diagonal (models/layers.py:55-56).
template <int PRE>
__device__ __forceinline__ float pre_value(float v, int i, int j) {
    if (PRE) return (i == j) ? 1.0f : fabsf(v);
    return v;
}

// Stage src[j, c0 + cc] (all rows j < n, cc < 32) into s[j * 33 + cc].
template <int PRE>
__device__ __forceinline__ void load_col_strip(const float* __restrict__ src, int n, int c0,
                                               float* __restrict__ s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = c0 + lane;
    for (int j = warp; j < n; j += kEwWarps) {
        float v = 0.f;
        if (c < n) v = pre_value<PRE>(__ldg(src + (size_t)j * n + c), j, c);
        s[j * kStripPad + lane] = v;
    }
}

// ---------------------------------------------------------------- normalise fwd
// utils/preprocessing.py:16-25
template <int PRE>
__global__ void __launch_bounds__(kEwThreads)
normalize_fwd_kernel(const float* __restrict__ mx, float* __restrict__ out,
                     float* __restrict__ r_out, float* __restrict__ m_out, int n) {
    extern __shared__ float smem[];
    float* r = smem;                 // [n]
    float* strip = smem + ((n + 31) & ~31);   // [n][33]
    const size_t base = (size_t)blockIdx.x * n * n;
    mx += base; out += base;
    if (m_out) m_out += base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // phase 1: rowsum -> r = rowsum^-1/2, inf -> 0  (:17-19)
    for (int i = warp; i < n; i += kEwWarps) {
        float acc = 0.f;
        for (int j = lane; j < n; j += 32) {
            float v = pre_value<PRE>(__ldg(mx + (size_t)i * n + j), i, j);
            if (m_out) m_out[(size_t)i * n + j] = v;
            acc += v;
        }
        acc = warp_sum(acc);
        float ri = 1.0f / sqrtf(acc);          // pow(rowsum, -0.5)
        if (isinf(ri)) ri = 0.f;
        if (lane == 0) { r[i] = ri; r_out[(size_t)blockIdx.x * n + i] = ri; }
    }
    __syncthreads();
    // phase 2: out[i,j] = (M[j,i] * r_i) * r_j   (:20-24: (mx @ D)^T @ D)
    for (int i0 = 0; i0 < n; i0 += 32) {
        load_col_strip<PRE>(mx, n, i0, strip);
        __syncthreads();
        const int rows = min(32, n - i0);
        for (int ii = warp; ii < rows; ii += kEwWarps) {
            const int i = i0 + ii;
            const float ri = r[i];
            float* dst = out + (size_t)i * n;
#pragma unroll 4
            for (int j = lane; j < n; j += 32) dst[j] = (strip[j * kStripPad + ii] * ri) * r[j];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- normalise bwd
// out[i,j] = r_i r_j M[j,i],  r_k = rho_k^-1/2 (0 where rho_k == 0),  T = grad_out
//   dr_k   = sum_j r_j (T[k,j] M[j,k] + T[j,k] M[k,j])
//   drho_k = -1/2 r_k^3 dr_k
//   dM[a,b] = T[b,a] r_a r_b + drho_a ;  PRE: dmx = dM * sgn(mx), 0 on the diagonal
template <int PRE>
__global__ void __launch_bounds__(kEwThreads)
normalize_bwd_kernel(const float* __restrict__ mx, const float* __restrict__ r_in,
                     const float* __restrict__ T, float* __restrict__ dmx, int n) {
    extern __shared__ float smem[];
    const int npad = (n + 31) & ~31;
    float* r = smem;                     // [npad]
    float* dr = r + npad;                // [npad]  row-term, later drho
    float* colpart = dr + npad;          // [kEwWarps][npad]
    float* strip = colpart + kEwWarps * npad;   // [n][33]
    const size_t base = (size_t)blockIdx.x * n * n;
    mx += base; T += base; dmx += base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < npad; i += kEwThreads) {
        r[i] = (i < n) ? r_in[(size_t)blockIdx.x * n + i] : 0.f;
        dr[i] = 0.f;
    }
    for (int i = threadIdx.x; i < kEwWarps * npad; i += kEwThreads) colpart[i] = 0.f;
    __syncthreads();

    // phase A: E[a,b] = T[a,b] * M[b,a];  dr_a += sum_b E[a,b] r_b ;  dr_b += sum_a E[a,b] r_a
    for (int a0 = 0; a0 < n; a0 += 32) {
        load_col_strip<PRE>(mx, n, a0, strip);       // strip[b][aa] = M[b, a0+aa]
        __syncthreads();
        const int rows = min(32, n - a0);
        for (int aa = warp; aa < rows; aa += kEwWarps) {
            const int a = a0 + aa;
            const float ra = r[a];
            float rowacc = 0.f;
            for (int b = lane; b < n; b += 32) {
                const float e = __ldg(T + (size_t)a * n + b) * strip[b * kStripPad + aa];
                rowacc += e * r[b];
                colpart[warp * npad + b] += e * ra;      // (warp, b) owned by exactly one lane
            }
            rowacc = warp_sum(rowacc);
            if (lane == 0) dr[a] = rowacc;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < n; k += kEwThreads) {
        float acc = dr[k];
#pragma unroll
        for (int w = 0; w < kEwWarps; ++w) acc += colpart[w * npad + k];
        const float rk = r[k];
        dr[k] = -0.5f * rk * rk * rk * acc;               // drho_k
    }
    __syncthreads();
    // phase B: dM[a,b] = T[b,a] r_a r_b + drho_a
    for (int a0 = 0; a0 < n; a0 += 32) {
        load_col_strip<0>(T, n, a0, strip);               // strip[b][aa] = T[b, a0+aa]
        __syncthreads();
        const int rows = min(32, n - a0);
        for (int aa = warp; aa < rows; aa += kEwWarps) {
            const int a = a0 + aa;
            const float ra = r[a], da = dr[a];
            const float* mrow = mx + (size_t)a * n;
            float* drow = dmx + (size_t)a * n;
#pragma unroll 4
            for (int b = lane; b < n; b += 32) {
                float g = (strip[b * kStripPad + aa] * ra) * r[b] + da;
                if (PRE) g = (a == b) ? 0.f : g * sgnf(__ldg(mrow + b));
                drow[b] = g;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- sym / diag / abs
// models/layers.py:60-63, models/gsr_model.py:32-37
__global__ void __launch_bounds__(kEwThreads)
sym_diag_abs_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
    extern __shared__ float strip[];
    const size_t base = (size_t)blockIdx.x * n * n;
    x += base; out += base;
    for (int i0 = 0; i0 < n; i0 += 32) {
        load_col_strip<0>(x, n, i0, strip);               // strip[j][ii] = x[j, i0+ii]
        __syncthreads();
        const int rows = min(32, n - i0);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int ii = warp; ii < rows; ii += kEwWarps) {
            const int i = i0 + ii;
            const float* xrow = x + (size_t)i * n;
            float* orow = out + (size_t)i * n;
#pragma unroll 4
            for (int j = lane; j < n; j += 32) {
                const float z = (__ldg(xrow + j) + strip[j * kStripPad + ii]) / 2.0f;
                orow[j] = (i == j) ? 1.0f : fabsf(z);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kEwThreads)
sym_diag_abs_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                        float* __restrict__ dx, int n) {
    extern __shared__ float smem[];
    float* sx = smem;                       // x[j, i0+ii]
    float* sg = smem + (size_t)n * kStripPad;   // g[j, i0+ii]
    const size_t base = (size_t)blockIdx.x * n * n;
    x += base; g += base; dx += base;
    for (int i0 = 0; i0 < n; i0 += 32) {
        load_col_strip<0>(x, n, i0, sx);
        load_col_strip<0>(g, n, i0, sg);
        __syncthreads();
        const int rows = min(32, n - i0);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int ii = warp; ii < rows; ii += kEwWarps) {
            const int i = i0 + ii;
            const float* xrow = x + (size_t)i * n;
            const float* grow = g + (size_t)i * n;
            float* drow = dx + (size_t)i * n;
#pragma unroll 4
            for (int j = lane; j < n; j += 32) {
                const float z = (__ldg(xrow + j) + sx[j * kStripPad + ii]) / 2.0f;
                const float s = sgnf(z);
                const float v = (__ldg(grow + j) * s + sg[j * kStripPad + ii] * s) / 2.0f;
                drow[j] = (i == j) ? 0.f : v;
            }
        }
        __syncthreads();
    }
}

__global__ void abs_diag1_kernel(const float* __restrict__ x, float* __restrict__ out,
                                 long long total, int n) {
    const long long nn = (long long)n * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long w = e % nn;
        const int i = (int)(w / n), j = (int)(w - (long long)i * n);
        out[e] = (i == j) ? 1.0f : fabsf(x[e]);
    }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return GSR_E_UNSUPPORTED;
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_normalize_adj_fwd(const float* mx, float* out, float* r_out, float* m_out,
                                     int batch, int n, int pre, void* stream) {
    if (!mx || !out || !r_out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "normalize_adj_fwd: bad argument");
    if (batch == 0) return 0;
    const size_t smem = ((size_t)((n + 31) & ~31) + (size_t)n * kStripPad) * sizeof(float);
    auto k = pre ? normalize_fwd_kernel<1> : normalize_fwd_kernel<0>;
    int rc = set_smem(k, smem);
    if (rc) return fail(rc, "normalize_adj_fwd: n too large for one CTA");
    k<<<batch, kEwThreads, smem, (cudaStream_t)stream>>>(mx, out, r_out, m_out, n);
    return launched("normalize_adj_fwd");
}

extern "C" int gsr_normalize_adj_bwd(const float* mx, const float* r, const float* grad_out,
                                     float* grad_mx, int batch, int n, int pre, void* stream) {
    if (!mx || !r || !grad_out || !grad_mx || batch < 0 || n <= 0) return fail(GSR_E_ARG, "normalize_adj_bwd: bad argument");
    if (batch == 0) return 0;
    const size_t npad = (n + 31) & ~31;
    const size_t smem = (npad * (2 + kEwWarps) + (size_t)n * kStripPad) * sizeof(float);
    auto k = pre ? normalize_bwd_kernel<1> : normalize_bwd_kernel<0>;
    int rc = set_smem(k, smem);
    if (rc) return fail(rc, "normalize_adj_bwd: n too large for one CTA");
    k<<<batch, kEwThreads, smem, (cudaStream_t)stream>>>(mx, r, grad_out, grad_mx, n);
    return launched("normalize_adj_bwd");
}

extern "C" int gsr_sym_diag_abs_fwd(const float* x, float* out, int batch, int n, void* stream) {
    if (!x || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "sym_diag_abs_fwd: bad argument");
    if (batch == 0) return 0;
    const size_t smem = (size_t)n * kStripPad * sizeof(float);
    int rc = set_smem(sym_diag_abs_fwd_kernel, smem);
    if (rc) return fail(rc, "sym_diag_abs_fwd: n too large for one CTA");
    sym_diag_abs_fwd_kernel<<<batch, kEwThreads, smem, (cudaStream_t)stream>>>(x, out, n);
    return launched("sym_diag_abs_fwd");
}

extern "C" int gsr_sym_diag_abs_bwd(const float* x, const float* grad_out, float* grad_x,
                                    int batch, int n, void* stream) {
    if (!x || !grad_out || !grad_x || batch < 0 || n <= 0) return fail(GSR_E_ARG, "sym_diag_abs_bwd: bad argument");
    if (batch == 0) return 0;
    const size_t smem = 2 * (size_t)n * kStripPad * sizeof(float);
    int rc = set_smem(sym_diag_abs_bwd_kernel, smem);
    if (rc) return fail(rc, "sym_diag_abs_bwd: n too large for one CTA");
    sym_diag_abs_bwd_kernel<<<batch, kEwThreads, smem, (cudaStream_t)stream>>>(x, grad_out, grad_x, n);
    return launched("sym_diag_abs_bwd");
}

extern "C" int gsr_abs_diag1_fwd(const float* x, float* out, int batch, int n, void* stream) {
    if (!x || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "abs_diag1_fwd: bad argument");
    if (batch == 0) return 0;
    const long long total = (long long)batch * n * n;
    const int blocks = (int)min((long long)kNumSMs * 8, (total + 255) / 256);
    abs_diag1_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, total, n);
    return launched("abs_diag1_fwd");
}
