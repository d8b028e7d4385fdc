// Memory-bound n x n kernels of the GSR-Net path: adjacency normalisation (fwd/bwd,
// optionally fused with abs + unit diagonal), symmetrise + unit diagonal + abs
// (fwd/bwd).  One CTA per subject matrix.  Every kernel walks the matrix in
// 32-row strips: the rows of the strip are read straight from global memory
// (coalesced), the matching 32-column strip (the transposed operand) is staged
// through padded shared memory so that both the global reads and the global
// writes stay coalesced.
#include "common.cuh"
#include <cooperative_groups.h>

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

// ---------------------------------------------------------------- sym / diag / abs, tile pairs
// The outputs are symmetric functions of (x_ij, x_ji) (and (g_ij, g_ji)), so the natural unit of
// work is the PAIR of 64 x 64 tiles (I, J), (J, I), I <= J: one 256-thread CTA loads both tiles
// with float4 row segments (all loads of the CTA are issued before the first use), combines tile
// (I, J) in registers against the transposed partner tile in padded shared memory, writes it, and
// writes the mirror tile from a transposed read of the result.  15 pairs per 268 x 268 matrix
// x 1024 subjects: enough independent CTAs (several per SM) to keep HBM busy, where the
// one-CTA-per-matrix strip kernels above are bound by their own barriers (34-44 % of HBM).
constexpr int kNormCluster = 16;        // normalise backward in one pass: CTAs (tile pairs) per subject cluster
int g_norm_bwd_cluster = 0;             // gsr_set_tuning("normalize_bwd_cluster", 0|1): three launches (default) / the one-pass cluster kernel for n <= 320
constexpr int kTS = 64;                 // tile edge
// Shared-memory image of a 64 x 64 tile: element (r, c) at tix(r, c) = 67 r + 2 (r >> 5) + c + 2 (c >> 5).  A warp touches
// a tile in two ways -- 16 lanes x 4 consecutive columns of one row (and the next row for the other 16 lanes) when a
// float4 row segment is stored, and 16 rows 4 apart at one column (and the next column) when the transposed partner is
// read.  With the plain 65-float pitch of the first version both patterns hit every bank twice (4 tx and 4 (tx + 8) are
// 32 apart).  The odd pitch keeps 4 P tx distinct for tx < 8, the skew of 2 after 32 columns / 32 rows moves the lanes
// tx >= 8 to the banks = 2 (mod 4), and the other half-warp (next row / next column: an odd offset) takes the odd banks:
// both patterns are conflict free.
constexpr int kTP = kTS + 3;            // pitch (odd)
constexpr int kTileFloats = kTS * kTP + 4;
__device__ __forceinline__ int tix(int r, int c) { return r * kTP + 2 * (r >> 5) + c + 2 * (c >> 5); }
constexpr int kTileThreads = 256;       // thread (ty, tx): rows ty + 16 p (p < 4), columns 4 tx .. 4 tx + 3

__device__ __forceinline__ void tile_pair_of(int pair, int T, int& I, int& J) {
    I = 0;
    while (pair >= T - I) { pair -= T - I; ++I; }       // row I of the upper triangle holds T - I pairs
    J = I + pair;
}
__device__ __forceinline__ float4 ld4_masked(const float* __restrict__ base, int n, int row, int col) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n && col < n) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * n + col));   // n % 4 == 0: all or nothing
    return v;
}
// the incoming gradient as the sum of two tensors (the second may be null): the sum is formed as it is read
__device__ __forceinline__ float4 ld4_masked2(const float* __restrict__ a, const float* __restrict__ b, int n, int row, int col) {
    float4 v = ld4_masked(a, n, row, col);
    if (b) {
        const float4 w = ld4_masked(b, n, row, col);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    return v;
}
__device__ __forceinline__ void st_tile(float* t, int r, int c, float4 v) {
    float* q = t + tix(r, c);               // c is a multiple of 4: the four columns share their 32-column group
    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
}

template <bool BWD>
__global__ void __launch_bounds__(kTileThreads, BWD ? 3 : 5)
sym_diag_abs_tiles_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ out, int n, int T) {
    extern __shared__ float smem[];
    float* bx = smem;                          // partner tile of x:  bx[jj][ii] = x[J0 + jj][I0 + ii]
    float* res = smem + kTileFloats;             // result tile (I, J)
    float* bg = res + kTileFloats;               // BWD: partner tile of g
    const size_t base = (size_t)blockIdx.y * n * n;
    x += base; out += base;
    if (BWD) g += base;
    int I, J;
    tile_pair_of(blockIdx.x, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 ax[4], pbx[4], ag[4], pbg[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int r = ty + 16 * p;
        ax[p] = ld4_masked(x, n, I0 + r, J0 + 4 * tx);
        pbx[p] = ld4_masked(x, n, J0 + r, I0 + 4 * tx);
        if (BWD) { ag[p] = ld4_masked(g, n, I0 + r, J0 + 4 * tx); pbg[p] = ld4_masked(g, n, J0 + r, I0 + 4 * tx); }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        st_tile(bx, ty + 16 * p, 4 * tx, pbx[p]);
        if (BWD) st_tile(bg, ty + 16 * p, 4 * tx, pbg[p]);
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ii = ty + 16 * p, i = I0 + ii;
        float a[4] = {ax[p].x, ax[p].y, ax[p].z, ax[p].w};
        float ga[4] = {0.f, 0.f, 0.f, 0.f};
        if (BWD) { ga[0] = ag[p].x; ga[1] = ag[p].y; ga[2] = ag[p].z; ga[3] = ag[p].w; }
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = 4 * tx + u, j = J0 + jj;
            const float z = (a[u] + bx[tix(jj, ii)]) / 2.0f;                  // (x + x^T) / 2
            if (!BWD) {
                o[u] = (i == j) ? 1.0f : fabsf(z);
            } else {
                const float sg = sgnf(z);
                const float v = (ga[u] * sg + bg[tix(jj, ii)] * sg) / 2.0f;
                o[u] = (i == j) ? 0.f : v;
            }
        }
        const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
        if (i < n && J0 + 4 * tx < n) *reinterpret_cast<float4*>(out + (size_t)i * n + J0 + 4 * tx) = o4;
        st_tile(res, ii, 4 * tx, o4);
    }
    if (I == J) return;                                                          // the diagonal tile is its own mirror
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int jj = ty + 16 * p, j = J0 + jj;                                // row of the mirror tile (J, I)
        if (j < n && I0 + 4 * tx < n) {
            const float4 o4 = make_float4(res[tix(4 * tx, jj)], res[tix(4 * tx + 1, jj)],
                                          res[tix(4 * tx + 2, jj)], res[tix(4 * tx + 3, jj)]);
            *reinterpret_cast<float4*>(out + (size_t)j * n + I0 + 4 * tx) = o4;
        }
    }
}

// ---------------------------------------------------------------- sym / diag / abs + L1 against a target, tile pairs
// The last two operations of a training step (models/gsr_model.py:32-37, then nn.L1Loss against the HR target,
// utils/gsr_train.py:39-43) as ONE pass: z = |(x + x^T) / 2| with a unit diagonal is formed in registers and
// compared with the target tile and its mirror; the prediction is written only if the caller asks for it.
//   FWD  partial[b][pair] = sum over the pair's elements of |z - y|  (double; folded in fixed order by
//        sym_l1_final_kernel)                                  -- reads x and y once, writes nothing of size n^2
//   BWD  dx = backward of sym-diag-abs applied to the L1 gradient sgn(z - y) * grad_loss[b] / n^2, recomputed from
//        x and y instead of being read from a [B, n, n] gradient tensor: the same expressions in the same order as
//        l1_bwd_kernel followed by sym_diag_abs_tiles_kernel<true>, so the result is bit identical to the two launches
template <bool BWD>
__global__ void __launch_bounds__(kTileThreads, BWD ? 3 : 5)
sym_diag_abs_l1_tiles_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ grad_loss,
                             float* __restrict__ out, double* __restrict__ partials, int n, int T) {
    extern __shared__ float smem[];
    float* bx = smem;                          // partner tile of x:  bx[jj][ii] = x[J0 + jj][I0 + ii]
    float* by = smem + kTileFloats;              // partner tile of y
    float* res = by + kTileFloats;               // result tile (I, J) (BWD, or FWD with a prediction to write)
    __shared__ double s_red[kTileThreads / 32];
    const size_t base = (size_t)blockIdx.y * n * n;
    x += base; y += base;
    if (out) out += base;
    int I, J;
    tile_pair_of(blockIdx.x, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 ax[4], pbx[4], ay[4], pby[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int r = ty + 16 * p;
        ax[p] = ld4_masked(x, n, I0 + r, J0 + 4 * tx);
        pbx[p] = ld4_masked(x, n, J0 + r, I0 + 4 * tx);
        ay[p] = ld4_masked(y, n, I0 + r, J0 + 4 * tx);
        pby[p] = ld4_masked(y, n, J0 + r, I0 + 4 * tx);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        st_tile(bx, ty + 16 * p, 4 * tx, pbx[p]);
        st_tile(by, ty + 16 * p, 4 * tx, pby[p]);
    }
    __syncthreads();
    const float scale = BWD ? grad_loss[blockIdx.y] / (float)((long long)n * n) : 0.f;
    float acc = 0.f;                           // FWD: <= 32 terms per thread in fp32, fp64 across threads
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ii = ty + 16 * p, i = I0 + ii;
        const float a[4] = {ax[p].x, ax[p].y, ax[p].z, ax[p].w};
        const float ya[4] = {ay[p].x, ay[p].y, ay[p].z, ay[p].w};
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = 4 * tx + u, j = J0 + jj;
            const float z = (a[u] + bx[tix(jj, ii)]) / 2.0f;                  // (x + x^T) / 2
            const float pred = (i == j) ? 1.0f : fabsf(z);
            const float yb = by[tix(jj, ii)];                                 // y[j][i]
            if (!BWD) {
                o[u] = pred;
                if (i < n && j < n) {
                    acc += fabsf(pred - ya[u]);
                    if (I != J) acc += fabsf(pred - yb);                        // the mirror element
                }
            } else {
                const float sg = sgnf(z);
                const float ga = sgnf(pred - ya[u]) * scale, gb = sgnf(pred - yb) * scale;
                const float v = (ga * sg + gb * sg) / 2.0f;
                o[u] = (i == j) ? 0.f : v;
            }
        }
        if (BWD || out) {
            const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
            if (i < n && J0 + 4 * tx < n) *reinterpret_cast<float4*>(out + (size_t)i * n + J0 + 4 * tx) = o4;
            st_tile(res, ii, 4 * tx, o4);
        }
    }
    if (!BWD) {
        double v = (double)acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kTileThreads / 32; ++w) t += s_red[w];
            partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
        if (!out) return;
    }
    if (I == J) return;                                                          // the diagonal tile is its own mirror
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int jj = ty + 16 * p, j = J0 + jj;                                // row of the mirror tile (J, I)
        if (j < n && I0 + 4 * tx < n) {
            const float4 o4 = make_float4(res[tix(4 * tx, jj)], res[tix(4 * tx + 1, jj)],
                                          res[tix(4 * tx + 2, jj)], res[tix(4 * tx + 3, jj)]);
            *reinterpret_cast<float4*>(out + (size_t)j * n + I0 + 4 * tx) = o4;
        }
    }
}

// loss[b] = (sum of the subject's pair partials, fixed order) / n^2
__global__ void sym_l1_final_kernel(const double* __restrict__ partials, float* __restrict__ loss, int pairs, long long count, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double t = 0.0;
    for (int c = 0; c < pairs; ++c) t += partials[(size_t)b * pairs + c];
    loss[b] = (float)(t / (double)count);
}

// ---------------------------------------------------------------- normalise, tile pairs
// Same decomposition as sym_diag_abs_tiles_kernel for utils/preprocessing.py:16-25 and its autograd:
//   fwd   rowsum kernel (warp per row, float4) -> r;  tile-pair kernel: out(I,J) = r_i r_j M(J,I)^T and its mirror
//   bwd   pass 1 (tile pairs): S = T(I,J) .* M(J,I)^T + T(J,I)^T .* M(I,J); partial sums S r over the tile go to
//         a [batch][tiles][n] workspace slot written by exactly one CTA (deterministic); a small kernel folds the
//         slots into drho;  pass 2 (tile pairs): dM(I,J) = r_i r_j T(J,I)^T + drho_i (x sgn(mx), zero diagonal if PRE)
template <int PRE>
__global__ void __launch_bounds__(256)
normalize_rowsum_kernel(const float* __restrict__ mx, float* __restrict__ r_out, float* __restrict__ m_out, int n) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const size_t base = (size_t)blockIdx.y * n * n + (size_t)i * n;
    float acc = 0.f;
    for (int c = lane; c < (n >> 2); c += 32) {
        float4 v = __ldg(reinterpret_cast<const float4*>(mx + base) + c);
        v.x = pre_value<PRE>(v.x, i, 4 * c); v.y = pre_value<PRE>(v.y, i, 4 * c + 1);
        v.z = pre_value<PRE>(v.z, i, 4 * c + 2); v.w = pre_value<PRE>(v.w, i, 4 * c + 3);
        if (m_out) reinterpret_cast<float4*>(m_out + base)[c] = v;
        acc += (v.x + v.y) + (v.z + v.w);
    }
    acc = warp_sum(acc);
    float ri = 1.0f / sqrtf(acc);          // pow(rowsum, -0.5)
    if (isinf(ri)) ri = 0.f;
    if (lane == 0) r_out[(size_t)blockIdx.y * n + i] = ri;
}

template <int PRE>
__device__ __forceinline__ float4 ld4_pre(const float* __restrict__ base, int n, int row, int col) {
    float4 v = ld4_masked(base, n, row, col);
    if (PRE) {
        v.x = pre_value<1>(v.x, row, col); v.y = pre_value<1>(v.y, row, col + 1);
        v.z = pre_value<1>(v.z, row, col + 2); v.w = pre_value<1>(v.w, row, col + 3);
        if (row >= n || col >= n) v = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return v;
}

template <int PRE>
__global__ void __launch_bounds__(kTileThreads)
normalize_scale_tiles_kernel(const float* __restrict__ mx, const float* __restrict__ r, float* __restrict__ out, int n, int T) {
    extern __shared__ float smem[];
    float* ta = smem;                          // ta[ii][jj] = M[I0 + ii][J0 + jj]
    float* tb = smem + kTileFloats;              // tb[jj][ii] = M[J0 + jj][I0 + ii]
    float* rI = tb + kTileFloats;                // [64]
    float* rJ = rI + kTS;                      // [64]
    const size_t base = (size_t)blockIdx.y * n * n;
    mx += base; out += base;
    r += (size_t)blockIdx.y * n;
    int I, J;
    tile_pair_of(blockIdx.x, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 va[4], vb[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        va[p] = ld4_pre<PRE>(mx, n, I0 + ty + 16 * p, J0 + 4 * tx);
        vb[p] = ld4_pre<PRE>(mx, n, J0 + ty + 16 * p, I0 + 4 * tx);
    }
    if (threadIdx.x < kTS) rI[threadIdx.x] = (I0 + threadIdx.x < n) ? r[I0 + threadIdx.x] : 0.f;
    else if (threadIdx.x < 2 * kTS) rJ[threadIdx.x - kTS] = (J0 + threadIdx.x - kTS < n) ? r[J0 + threadIdx.x - kTS] : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) { st_tile(ta, ty + 16 * p, 4 * tx, va[p]); st_tile(tb, ty + 16 * p, 4 * tx, vb[p]); }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ii = ty + 16 * p, i = I0 + ii;
        if (i < n && J0 + 4 * tx < n) {                          // out[i][j] = (M[j][i] r_i) r_j
            const float ri = rI[ii];
            float4 o;
            o.x = (tb[tix(4 * tx, ii)] * ri) * rJ[4 * tx];
            o.y = (tb[tix(4 * tx + 1, ii)] * ri) * rJ[4 * tx + 1];
            o.z = (tb[tix(4 * tx + 2, ii)] * ri) * rJ[4 * tx + 2];
            o.w = (tb[tix(4 * tx + 3, ii)] * ri) * rJ[4 * tx + 3];
            *reinterpret_cast<float4*>(out + (size_t)i * n + J0 + 4 * tx) = o;
        }
        if (I != J) {
            const int jj = ty + 16 * p, j = J0 + jj;
            if (j < n && I0 + 4 * tx < n) {                      // mirror tile: out[j][i] = (M[i][j] r_j) r_i
                const float rj = rJ[jj];
                float4 o;
                o.x = (ta[tix(4 * tx, jj)] * rj) * rI[4 * tx];
                o.y = (ta[tix(4 * tx + 1, jj)] * rj) * rI[4 * tx + 1];
                o.z = (ta[tix(4 * tx + 2, jj)] * rj) * rI[4 * tx + 2];
                o.w = (ta[tix(4 * tx + 3, jj)] * rj) * rI[4 * tx + 3];
                *reinterpret_cast<float4*>(out + (size_t)j * n + I0 + 4 * tx) = o;
            }
        }
    }
}

// bwd pass 1: partial dr over one tile pair.  part[b][slot][row]: rows of I from the pair (I, J) go to slot J,
// rows of J to slot I (every (slot, row block) is written by exactly one CTA).
template <int PRE>
__global__ void __launch_bounds__(kTileThreads, 3)
normalize_bwd_partial_tiles_kernel(const float* __restrict__ mx, const float* __restrict__ r, const float* __restrict__ Tg,
                                   const float* __restrict__ Tg2, float* __restrict__ part, int n, int T) {
    extern __shared__ float smem[];
    float* mb = smem;                          // M(J,I):  mb[jj][ii]
    float* tb = smem + kTileFloats;              // T(J,I):  tb[jj][ii]
    float* sS = tb + kTileFloats;                // S[ii][jj]
    float* rI = sS + kTileFloats;
    float* rJ = rI + kTS;
    const size_t base = (size_t)blockIdx.y * n * n;
    mx += base; Tg += base;
    if (Tg2) Tg2 += base;
    r += (size_t)blockIdx.y * n;
    part += (size_t)blockIdx.y * T * n;
    int I, J;
    tile_pair_of(blockIdx.x, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 ma[4], vmb[4], tA[4], vtb[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        ma[p] = ld4_pre<PRE>(mx, n, I0 + ty + 16 * p, J0 + 4 * tx);
        vmb[p] = ld4_pre<PRE>(mx, n, J0 + ty + 16 * p, I0 + 4 * tx);
        tA[p] = ld4_masked2(Tg, Tg2, n, I0 + ty + 16 * p, J0 + 4 * tx);
        vtb[p] = ld4_masked2(Tg, Tg2, n, J0 + ty + 16 * p, I0 + 4 * tx);
    }
    if (threadIdx.x < kTS) rI[threadIdx.x] = (I0 + threadIdx.x < n) ? r[I0 + threadIdx.x] : 0.f;
    else if (threadIdx.x < 2 * kTS) rJ[threadIdx.x - kTS] = (J0 + threadIdx.x - kTS < n) ? r[J0 + threadIdx.x - kTS] : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) { st_tile(mb, ty + 16 * p, 4 * tx, vmb[p]); st_tile(tb, ty + 16 * p, 4 * tx, vtb[p]); }
    __syncthreads();
    // S[ii][jj] = T[i][j] M[j][i] + T[j][i] M[i][j]   (on the diagonal tile the two terms are the (i,j) and (j,i) of one sum)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ii = ty + 16 * p;
        const float a4[4] = {ma[p].x, ma[p].y, ma[p].z, ma[p].w};
        const float t4[4] = {tA[p].x, tA[p].y, tA[p].z, tA[p].w};
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = 4 * tx + u;
            op[u] = t4[u] * mb[tix(jj, ii)] + tb[tix(jj, ii)] * a4[u];
        }
        st_tile(sS, ii, 4 * tx, o);
    }
    __syncthreads();
    // row sums (rows of I, weights r_J) by threads 0..63, column sums (rows of J, weights r_I) by threads 64..127
    if (threadIdx.x < kTS) {
        const int ii = threadIdx.x;
        if (I0 + ii < n) {
            float acc = 0.f;
            for (int jj = 0; jj < kTS; ++jj) acc = fmaf(sS[tix(ii, jj)], rJ[jj], acc);
            part[(size_t)J * n + I0 + ii] = acc;
        }
    } else if (threadIdx.x < 2 * kTS && I != J) {
        const int jj = threadIdx.x - kTS;
        if (J0 + jj < n) {
            float acc = 0.f;
            for (int ii = 0; ii < kTS; ++ii) acc = fmaf(sS[tix(ii, jj)], rI[ii], acc);
            part[(size_t)I * n + J0 + jj] = acc;
        }
    }
}

// drho_k = -1/2 r_k^3 sum_slots part[slot][k]   (fixed order)
__global__ void normalize_bwd_drho_kernel(const float* __restrict__ part, const float* __restrict__ r, float* __restrict__ drho,
                                          int n, int T, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long b = e / n;
    const int k = (int)(e - b * n);
    float acc = 0.f;
    for (int sl = 0; sl < T; ++sl) acc += part[((size_t)b * T + sl) * n + k];
    const float rk = r[e];
    drho[e] = -0.5f * rk * rk * rk * acc;
}

// bwd pass 2: dM(I,J)[i][j] = (T[j][i] r_i) r_j + drho_i;  PRE: x sgn(mx[i][j]), zero diagonal
template <int PRE>
__global__ void __launch_bounds__(kTileThreads, 3)
normalize_bwd_apply_tiles_kernel(const float* __restrict__ mx, const float* __restrict__ r, const float* __restrict__ drho,
                                 const float* __restrict__ Tg, const float* __restrict__ Tg2, float* __restrict__ dmx, int n, int T) {
    extern __shared__ float smem[];
    float* ta = smem;                          // T(I,J): ta[ii][jj]
    float* tb = smem + kTileFloats;              // T(J,I): tb[jj][ii]
    float* rI = tb + kTileFloats;
    float* rJ = rI + kTS;
    float* dI = rJ + kTS;
    float* dJ = dI + kTS;
    const size_t base = (size_t)blockIdx.y * n * n;
    mx += base; Tg += base; dmx += base;
    if (Tg2) Tg2 += base;
    r += (size_t)blockIdx.y * n;
    drho += (size_t)blockIdx.y * n;
    int I, J;
    tile_pair_of(blockIdx.x, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 vta[4], vtb[4], xa[4], xb[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        vta[p] = ld4_masked2(Tg, Tg2, n, I0 + ty + 16 * p, J0 + 4 * tx);
        vtb[p] = ld4_masked2(Tg, Tg2, n, J0 + ty + 16 * p, I0 + 4 * tx);
        if (PRE) { xa[p] = ld4_masked(mx, n, I0 + ty + 16 * p, J0 + 4 * tx); xb[p] = ld4_masked(mx, n, J0 + ty + 16 * p, I0 + 4 * tx); }
    }
    if (threadIdx.x < kTS) {
        const int k = I0 + threadIdx.x;
        rI[threadIdx.x] = k < n ? r[k] : 0.f; dI[threadIdx.x] = k < n ? drho[k] : 0.f;
    } else if (threadIdx.x < 2 * kTS) {
        const int k = J0 + threadIdx.x - kTS;
        rJ[threadIdx.x - kTS] = k < n ? r[k] : 0.f; dJ[threadIdx.x - kTS] = k < n ? drho[k] : 0.f;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) { st_tile(ta, ty + 16 * p, 4 * tx, vta[p]); st_tile(tb, ty + 16 * p, 4 * tx, vtb[p]); }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ii = ty + 16 * p, i = I0 + ii;
        if (i < n && J0 + 4 * tx < n) {
            const float ri = rI[ii], di = dI[ii];
            const float xs[4] = {PRE ? xa[p].x : 0.f, PRE ? xa[p].y : 0.f, PRE ? xa[p].z : 0.f, PRE ? xa[p].w : 0.f};
            float4 o;
            float* op = &o.x;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = 4 * tx + u;
                float gv = (tb[tix(jj, ii)] * ri) * rJ[jj] + di;
                if (PRE) gv = (i == J0 + jj) ? 0.f : gv * sgnf(xs[u]);
                op[u] = gv;
            }
            *reinterpret_cast<float4*>(dmx + (size_t)i * n + J0 + 4 * tx) = o;
        }
        if (I != J) {
            const int jj = ty + 16 * p, j = J0 + jj;
            if (j < n && I0 + 4 * tx < n) {
                const float rj = rJ[jj], dj = dJ[jj];
                const float xs[4] = {PRE ? xb[p].x : 0.f, PRE ? xb[p].y : 0.f, PRE ? xb[p].z : 0.f, PRE ? xb[p].w : 0.f};
                float4 o;
                float* op = &o.x;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ii2 = 4 * tx + u;
                    float gv = (ta[tix(ii2, jj)] * rj) * rI[ii2] + dj;
                    if (PRE) gv = (j == I0 + ii2) ? 0.f : gv * sgnf(xs[u]);
                    op[u] = gv;
                }
                *reinterpret_cast<float4*>(dmx + (size_t)j * n + I0 + 4 * tx) = o;
            }
        }
    }
}

// bwd as ONE pass (n <= 320: at most 15 tile pairs): the CTAs of a subject form a thread-block CLUSTER of 16, so they
// are resident together and can exchange their partial sums through distributed shared memory between the two halves of
// the algorithm -- every CTA keeps its pair of tiles on chip (T(I,J) + T(J,I) in shared memory, the signs of mx in two
// registers) instead of reading T, T2 and mx a second time: 4 tensors of traffic (T, T2, mx in, dM out) instead of 7.
// Same expressions and the same summation order as the three launches above (partial -> drho -> apply): bit identical.
// MEASURED SLOWER and therefore off by default (1024 x 268, two incoming gradients: 611 us at 3 CTAs per SM, 770 us at 2,
// against 550 us for the three launches): a cluster of 16 holds 16 CTA slots of one GPC from its first load to its last
// store, every CTA waits twice for the slowest of the 16, and load, exchange and store phases of a CTA do not overlap;
// the three launches keep 3 independent CTAs per SM streaming.  Kept as a switch with its parity test.
__device__ __forceinline__ int tile_pair_rank(int I, int J, int T) { return I * T - I * (I - 1) / 2 + (J - I); }     // inverse of tile_pair_of, I <= J
__device__ __forceinline__ uint32_t sign_code(float v) { return v > 0.f ? 1u : (v < 0.f ? 2u : 0u); }               // 2 bits: sgnf(v) = 0, +1, -1
__device__ __forceinline__ float sign_of_code(uint32_t c) { return c == 1u ? 1.0f : (c == 2u ? -1.0f : 0.0f); }

template <int PRE>
__global__ void __launch_bounds__(kTileThreads, 3)
normalize_bwd_cluster_kernel(const float* __restrict__ mx, const float* __restrict__ r, const float* __restrict__ Tg,
                             const float* __restrict__ Tg2, float* __restrict__ dmx, int n, int T) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float smem[];
    float* mb = smem;                          // M(J,I):  mb[jj][ii]
    float* tb = smem + kTileFloats;              // T(J,I):  tb[jj][ii]
    float* sS = tb + kTileFloats;                // S[ii][jj], then T(I,J): ta[ii][jj]
    float* rI = sS + kTileFloats;
    float* rJ = rI + kTS;
    float* dI = rJ + kTS;
    float* dJ = dI + kTS;
    float* pout = dJ + kTS;                    // [2][64]: this pair's row part (rows of I, slot J) and column part (rows of J, slot I)
    const int pairs = T * (T + 1) / 2;
    const int rank = (int)blockIdx.x;          // = rank in the cluster (cluster = the grid's x extent)
    const bool live = rank < pairs;
    const size_t base = (size_t)blockIdx.y * n * n;
    mx += base; Tg += base; dmx += base;
    if (Tg2) Tg2 += base;
    r += (size_t)blockIdx.y * n;
    int I = 0, J = 0;
    if (live) tile_pair_of(rank, T, I, J);
    const int I0 = I * kTS, J0 = J * kTS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 tA[4];
    uint32_t sa = 0u, sb = 0u;                 // PRE: 2-bit signs of mx(I,J) / mx(J,I) at this thread's 16 + 16 positions
    if (live) {
        float4 ma[4], vmb[4], vtb[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            ma[p] = ld4_masked(mx, n, I0 + ty + 16 * p, J0 + 4 * tx);
            vmb[p] = ld4_masked(mx, n, J0 + ty + 16 * p, I0 + 4 * tx);
            tA[p] = ld4_masked2(Tg, Tg2, n, I0 + ty + 16 * p, J0 + 4 * tx);
            vtb[p] = ld4_masked2(Tg, Tg2, n, J0 + ty + 16 * p, I0 + 4 * tx);
        }
        if (threadIdx.x < kTS) rI[threadIdx.x] = (I0 + threadIdx.x < n) ? r[I0 + threadIdx.x] : 0.f;
        else if (threadIdx.x < 2 * kTS) rJ[threadIdx.x - kTS] = (J0 + threadIdx.x - kTS < n) ? r[J0 + threadIdx.x - kTS] : 0.f;
        if (PRE) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float a4[4] = {ma[p].x, ma[p].y, ma[p].z, ma[p].w}, b4[4] = {vmb[p].x, vmb[p].y, vmb[p].z, vmb[p].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) { sa |= sign_code(a4[u]) << (2 * (4 * p + u)); sb |= sign_code(b4[u]) << (2 * (4 * p + u)); }
                const int ra = I0 + ty + 16 * p, ca = J0 + 4 * tx, rb = J0 + ty + 16 * p, cb = I0 + 4 * tx;
                ma[p] = (ra >= n || ca >= n) ? make_float4(0.f, 0.f, 0.f, 0.f)
                        : make_float4(pre_value<1>(a4[0], ra, ca), pre_value<1>(a4[1], ra, ca + 1), pre_value<1>(a4[2], ra, ca + 2), pre_value<1>(a4[3], ra, ca + 3));
                vmb[p] = (rb >= n || cb >= n) ? make_float4(0.f, 0.f, 0.f, 0.f)
                         : make_float4(pre_value<1>(b4[0], rb, cb), pre_value<1>(b4[1], rb, cb + 1), pre_value<1>(b4[2], rb, cb + 2), pre_value<1>(b4[3], rb, cb + 3));
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) { st_tile(mb, ty + 16 * p, 4 * tx, vmb[p]); st_tile(tb, ty + 16 * p, 4 * tx, vtb[p]); }
        __syncthreads();
        // S[ii][jj] = T[i][j] M[j][i] + T[j][i] M[i][j]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ii = ty + 16 * p;
            const float a4[4] = {ma[p].x, ma[p].y, ma[p].z, ma[p].w};
            const float t4[4] = {tA[p].x, tA[p].y, tA[p].z, tA[p].w};
            float4 o;
            float* op = &o.x;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = 4 * tx + u;
                op[u] = t4[u] * mb[tix(jj, ii)] + tb[tix(jj, ii)] * a4[u];
            }
            st_tile(sS, ii, 4 * tx, o);
        }
        __syncthreads();
        if (threadIdx.x < kTS) {
            const int ii = threadIdx.x;
            float acc = 0.f;
            for (int jj = 0; jj < kTS; ++jj) acc = fmaf(sS[tix(ii, jj)], rJ[jj], acc);
            pout[ii] = acc;
        } else if (threadIdx.x < 2 * kTS) {
            const int jj = threadIdx.x - kTS;
            float acc = 0.f;
            if (I != J) for (int ii = 0; ii < kTS; ++ii) acc = fmaf(sS[tix(ii, jj)], rI[ii], acc);
            pout[kTS + jj] = acc;
        }
    }
    cluster.sync();                            // every pair's parts are in its CTA's shared memory
    if (live) {
        // drho_k = -1/2 r_k^3 sum over the slots sl = 0 .. T-1 (fixed order) of the part the pair {R, sl} holds for tile row R
        if (threadIdx.x < 2 * kTS) {
            const bool second = threadIdx.x >= kTS;
            const int R = second ? J : I, k = threadIdx.x - (second ? kTS : 0);
            float acc = 0.f;
            for (int sl = 0; sl < T; ++sl) {
                const int pi = min(R, sl), pj = max(R, sl);
                const float* remote = cluster.map_shared_rank(pout, tile_pair_rank(pi, pj, T));
                acc += remote[(R <= sl ? 0 : kTS) + k];       // rows of I (slot J = sl)  /  rows of J (slot I = sl)
            }
            const float rk = second ? rJ[k] : rI[k];
            (second ? dJ : dI)[k] = -0.5f * rk * rk * rk * acc;
        }
        // T(I,J) into shared memory for the mirror tile (S is consumed)
#pragma unroll
        for (int p = 0; p < 4; ++p) st_tile(sS, ty + 16 * p, 4 * tx, tA[p]);
    }
    cluster.barrier_arrive();                  // this CTA has read its peers' parts (the matching wait is at the end: nobody exits before)
    if (live) {
        __syncthreads();
        const float* ta = sS;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ii = ty + 16 * p, i = I0 + ii;
            if (i < n && J0 + 4 * tx < n) {
                const float ri = rI[ii], di = dI[ii];
                float4 o;
                float* op = &o.x;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = 4 * tx + u;
                    float gv = (tb[tix(jj, ii)] * ri) * rJ[jj] + di;
                    if (PRE) gv = (i == J0 + jj) ? 0.f : gv * sign_of_code((sa >> (2 * (4 * p + u))) & 3u);
                    op[u] = gv;
                }
                *reinterpret_cast<float4*>(dmx + (size_t)i * n + J0 + 4 * tx) = o;
            }
            if (I != J) {
                const int jj = ty + 16 * p, j = J0 + jj;
                if (j < n && I0 + 4 * tx < n) {
                    const float rj = rJ[jj], dj = dJ[jj];
                    float4 o;
                    float* op = &o.x;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ii2 = 4 * tx + u;
                        float gv = (ta[tix(ii2, jj)] * rj) * rI[ii2] + dj;
                        if (PRE) gv = (j == I0 + ii2) ? 0.f : gv * sign_of_code((sb >> (2 * (4 * p + u))) & 3u);
                        op[u] = gv;
                    }
                    *reinterpret_cast<float4*>(dmx + (size_t)j * n + I0 + 4 * tx) = o;
                }
            }
        }
    }
    cluster.barrier_wait();
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
    if (n % 4 == 0 && n >= 2 * kTS && aligned16(mx) && aligned16(out) && (!m_out || aligned16(m_out)) && batch <= 65535) {
        cudaStream_t st = (cudaStream_t)stream;
        const int T = (n + kTS - 1) / kTS;
        const size_t tsmem = (2 * kTileFloats + 2 * kTS) * sizeof(float);
        if (pre) normalize_rowsum_kernel<1><<<dim3((n + 7) / 8, batch), 256, 0, st>>>(mx, r_out, m_out, n);
        else     normalize_rowsum_kernel<0><<<dim3((n + 7) / 8, batch), 256, 0, st>>>(mx, r_out, m_out, n);
        int rc1 = launched("normalize_rowsum");
        if (rc1) return rc1;
        if (pre) normalize_scale_tiles_kernel<1><<<dim3(T * (T + 1) / 2, batch), kTileThreads, tsmem, st>>>(mx, r_out, out, n, T);
        else     normalize_scale_tiles_kernel<0><<<dim3(T * (T + 1) / 2, batch), kTileThreads, tsmem, st>>>(mx, r_out, out, n, T);
        return launched("normalize_adj_fwd");
    }
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
    if (n % 4 == 0 && aligned16(x) && aligned16(out) && batch <= 65535) {
        const int T = (n + kTS - 1) / kTS;
        sym_diag_abs_tiles_kernel<false><<<dim3(T * (T + 1) / 2, batch), kTileThreads, 2 * kTileFloats * sizeof(float), (cudaStream_t)stream>>>(x, nullptr, out, n, T);
        return launched("sym_diag_abs_fwd");
    }
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
    if (n % 4 == 0 && aligned16(x) && aligned16(grad_out) && aligned16(grad_x) && batch <= 65535) {
        const int T = (n + kTS - 1) / kTS;
        const size_t tsmem = 3 * kTileFloats * sizeof(float);
        int rc2 = set_smem(sym_diag_abs_tiles_kernel<true>, tsmem);
        if (rc2) return fail(rc2, "sym_diag_abs_bwd: cudaFuncSetAttribute failed");
        sym_diag_abs_tiles_kernel<true><<<dim3(T * (T + 1) / 2, batch), kTileThreads, tsmem, (cudaStream_t)stream>>>(x, grad_out, grad_x, n, T);
        return launched("sym_diag_abs_bwd");
    }
    const size_t smem = 2 * (size_t)n * kStripPad * sizeof(float);
    int rc = set_smem(sym_diag_abs_bwd_kernel, smem);
    if (rc) return fail(rc, "sym_diag_abs_bwd: n too large for one CTA");
    sym_diag_abs_bwd_kernel<<<batch, kEwThreads, smem, (cudaStream_t)stream>>>(x, grad_out, grad_x, n);
    return launched("sym_diag_abs_bwd");
}

extern "C" int gsr_sym_diag_abs_l1_partials(int n) {
    const int T = (n + kTS - 1) / kTS;
    return T * (T + 1) / 2;
}

extern "C" int gsr_sym_diag_abs_l1_fwd(const float* x, const float* y, float* pred_out, float* loss_out, double* partials,
                                       int batch, int n, void* stream) {
    if (!x || !y || !loss_out || !partials || batch < 0 || n <= 0) return fail(GSR_E_ARG, "sym_diag_abs_l1_fwd: bad argument");
    if (batch == 0) return 0;
    if (n % 4 != 0 || !aligned16(x) || !aligned16(y) || (pred_out && !aligned16(pred_out)) || batch > 65535)
        return fail(GSR_E_UNSUPPORTED, "sym_diag_abs_l1_fwd: needs n % 4 == 0, 16-byte aligned tensors and batch <= 65535 (use sym_diag_abs + l1_loss)");
    const int T = (n + kTS - 1) / kTS, pairs = T * (T + 1) / 2;
    const size_t tsmem = (pred_out ? 3 : 2) * kTileFloats * sizeof(float);      // the result tile only when a prediction is written
    int rc = set_smem(sym_diag_abs_l1_tiles_kernel<false>, tsmem);
    if (rc) return fail(rc, "sym_diag_abs_l1_fwd: cudaFuncSetAttribute failed");
    sym_diag_abs_l1_tiles_kernel<false><<<dim3(pairs, batch), kTileThreads, tsmem, (cudaStream_t)stream>>>(x, y, nullptr, pred_out, partials, n, T);
    if ((rc = launched("sym_diag_abs_l1_fwd"))) return rc;
    sym_l1_final_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partials, loss_out, pairs, (long long)n * n, batch);
    return launched("sym_l1_final");
}

extern "C" int gsr_sym_diag_abs_l1_bwd(const float* x, const float* y, const float* grad_loss, float* grad_x,
                                       int batch, int n, void* stream) {
    if (!x || !y || !grad_loss || !grad_x || batch < 0 || n <= 0) return fail(GSR_E_ARG, "sym_diag_abs_l1_bwd: bad argument");
    if (batch == 0) return 0;
    if (n % 4 != 0 || !aligned16(x) || !aligned16(y) || !aligned16(grad_x) || batch > 65535)
        return fail(GSR_E_UNSUPPORTED, "sym_diag_abs_l1_bwd: needs n % 4 == 0, 16-byte aligned tensors and batch <= 65535");
    const int T = (n + kTS - 1) / kTS;
    const size_t tsmem = 3 * kTileFloats * sizeof(float);
    int rc = set_smem(sym_diag_abs_l1_tiles_kernel<true>, tsmem);
    if (rc) return fail(rc, "sym_diag_abs_l1_bwd: cudaFuncSetAttribute failed");
    sym_diag_abs_l1_tiles_kernel<true><<<dim3(T * (T + 1) / 2, batch), kTileThreads, tsmem, (cudaStream_t)stream>>>(x, y, grad_loss, grad_x, nullptr, n, T);
    return launched("sym_diag_abs_l1_bwd");
}

extern "C" int gsr_abs_diag1_fwd(const float* x, float* out, int batch, int n, void* stream) {
    if (!x || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "abs_diag1_fwd: bad argument");
    if (batch == 0) return 0;
    const long long total = (long long)batch * n * n;
    const int blocks = (int)min((long long)kNumSMs * 8, (total + 255) / 256);
    abs_diag1_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, total, n);
    return launched("abs_diag1_fwd");
}

extern "C" size_t gsr_normalize_adj_bwd_workspace(int batch, int n) {
    if (batch <= 0 || n <= 0) return 0;
    const size_t T = (n + kTS - 1) / kTS;
    return (size_t)batch * (T + 1) * n;                     // floats: [batch][T][n] partial sums + [batch][n] drho
}

// grad_out2 (may be NULL): a second contribution to the incoming gradient, added to grad_out element by element as the two
// passes read it -- d adj = 2 dX0 adj (+) the graph convolutions' contribution (models/layers.py:57-59, models/gsr_model.py:30-32)
// without a GEMM addend or a separate addition pass.  No strip-kernel fallback: GSR_E_UNSUPPORTED when the tile path does not apply.
extern "C" int gsr_normalize_adj_bwd_tiles2(const float* mx, const float* r, const float* grad_out, const float* grad_out2, float* grad_mx,
                                            float* workspace, int batch, int n, int pre, void* stream) {
    if (!mx || !r || !grad_out || !grad_mx || batch < 0 || n <= 0) return fail(GSR_E_ARG, "normalize_adj_bwd_tiles: bad argument");
    if (batch == 0) return 0;
    if (!workspace || n % 4 != 0 || !aligned16(mx) || !aligned16(grad_out) || !aligned16(grad_mx) || batch > 65535 ||
        (grad_out2 && !aligned16(grad_out2))) {
        if (grad_out2) return fail(GSR_E_UNSUPPORTED, "normalize_adj_bwd_tiles2: two gradients need the tile path (workspace, n % 4 == 0, 16-byte aligned operands)");
        return gsr_normalize_adj_bwd(mx, r, grad_out, grad_mx, batch, n, pre, stream);      // strip kernel: any n, no workspace
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int T = (n + kTS - 1) / kTS;
    if (g_norm_bwd_cluster && T * (T + 1) / 2 <= kNormCluster) {
        // one pass: the tile pairs of a subject as a cluster of 16 CTAs (opt-in size); a refused launch falls back for good
        static bool refused = false;
        if (!refused) {
            const size_t smemc = (3 * kTileFloats + 6 * kTS) * sizeof(float);
            auto kern = pre ? normalize_bwd_cluster_kernel<1> : normalize_bwd_cluster_kernel<0>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e == cudaSuccess) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(kNormCluster, batch, 1);
                cfg.blockDim = dim3(kTileThreads, 1, 1);
                cfg.dynamicSmemBytes = smemc;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = kNormCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                e = cudaLaunchKernelEx(&cfg, kern, mx, r, grad_out, grad_out2, grad_mx, n, T);
            }
            if (e == cudaSuccess) return launched("normalize_adj_bwd_cluster");
            cudaGetLastError();
            refused = true;
        }
    }
    float* part = workspace;
    float* drho = workspace + (size_t)batch * T * n;
    const size_t smem1 = (3 * kTileFloats + 2 * kTS) * sizeof(float), smem2 = (2 * kTileFloats + 4 * kTS) * sizeof(float);
    int rc = pre ? set_smem(normalize_bwd_partial_tiles_kernel<1>, smem1) : set_smem(normalize_bwd_partial_tiles_kernel<0>, smem1);
    if (rc) return fail(rc, "normalize_adj_bwd_tiles: cudaFuncSetAttribute failed");
    const dim3 grid(T * (T + 1) / 2, batch);
    if (pre) normalize_bwd_partial_tiles_kernel<1><<<grid, kTileThreads, smem1, st>>>(mx, r, grad_out, grad_out2, part, n, T);
    else     normalize_bwd_partial_tiles_kernel<0><<<grid, kTileThreads, smem1, st>>>(mx, r, grad_out, grad_out2, part, n, T);
    if ((rc = launched("normalize_bwd_partial"))) return rc;
    const long long total = (long long)batch * n;
    normalize_bwd_drho_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(part, r, drho, n, T, total);
    if ((rc = launched("normalize_bwd_drho"))) return rc;
    if (pre) normalize_bwd_apply_tiles_kernel<1><<<grid, kTileThreads, smem2, st>>>(mx, r, drho, grad_out, grad_out2, grad_mx, n, T);
    else     normalize_bwd_apply_tiles_kernel<0><<<grid, kTileThreads, smem2, st>>>(mx, r, drho, grad_out, grad_out2, grad_mx, n, T);
    return launched("normalize_adj_bwd");
}

extern "C" int gsr_normalize_adj_bwd_tiles(const float* mx, const float* r, const float* grad_out, float* grad_mx,
                                           float* workspace, int batch, int n, int pre, void* stream) {
    return gsr_normalize_adj_bwd_tiles2(mx, r, grad_out, nullptr, grad_mx, workspace, batch, n, pre, stream);
}
