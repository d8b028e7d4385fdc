// MatrixVectorizer (utils/matrix_vectorizer.py) as index-math gather / scatter kernels.
//   vectorize      :24-58  strict upper triangle, COLUMN-major walk: (r, c), r < c, lands at
//                          start(c) + r with start(c) = c(c-1)/2; include_diagonal appends the
//                          SUB-diagonal element (c+1, c) after column c, so start(c) = c(c-1)/2 + c.
//   anti_vectorize :61-95  ROW-major strict upper fill (torch.triu_indices order), mirrored;
//                          include_diagonal puts vec[n_el + i] on the MAIN diagonal; fp32 output.
// Both are pure byte movement (bit exact).  vectorize reads 32x32 tiles row-wise (coalesced),
// transposes them through shared memory and writes runs of consecutive r (coalesced).
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ long long vec_start(int c, int diag) {
    return (long long)c * (c - 1) / 2 + (diag ? c : 0);
}

template <typename T>
__global__ void __launch_bounds__(256)
vectorize_kernel(const T* __restrict__ M, T* __restrict__ out, int n, long long len, int diag) {
    __shared__ T tile[32][33];
    const int ct = blockIdx.x, rt = blockIdx.y;
    if (rt > ct) return;                       // tile entirely below the diagonal
    const T* Mb = M + (size_t)blockIdx.z * n * n;
    T* ob = out + (size_t)blockIdx.z * len;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;      // 8 warps
    const int r0 = rt * 32, c0 = ct * 32;
    for (int rr = w; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + lane;
        if (r < n && c < n) tile[rr][lane] = Mb[(size_t)r * n + c];
    }
    __syncthreads();
    for (int cc = w; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + lane;
        if (c < n && r < c) ob[vec_start(c, diag) + r] = tile[lane][cc];
    }
    if (diag && rt == ct) {                    // sub-diagonal element (c+1, c) closes column c
        const int c = c0 + threadIdx.x;
        if (threadIdx.x < 32 && c + 1 < n) ob[vec_start(c, 1) + c] = Mb[(size_t)(c + 1) * n + c];
    }
}

__global__ void __launch_bounds__(256)
anti_vectorize_kernel(const float* __restrict__ vec, float* __restrict__ out, int n, long long vec_len, int diag) {
    const float* vb = vec + (size_t)blockIdx.y * vec_len;
    float* ob = out + (size_t)blockIdx.y * n * n;
    const long long n_el = (long long)n * (n - 1) / 2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * n;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), c = (int)(e - (long long)r * n);
        float v = 0.f;
        if (r != c) {
            const int a = min(r, c), b = max(r, c);
            v = vb[(long long)a * n - (long long)a * (a + 1) / 2 + (b - a - 1)];
        } else if (diag) {
            v = vb[n_el + r];
        }
        ob[e] = v;
    }
}

// training_gsr.ipynb:112-121 (submission): matrix[np.triu_indices(n, k=1)] -- the strict upper triangle in
// ROW-major order, the exact inverse of anti_vectorize's fill.  Element t of the output is (r, c) with
// t = r n - r (r + 1) / 2 + (c - r - 1); one thread per output element (coalesced writes, the reads of one
// matrix row are contiguous), r recovered from t with a float estimate fixed up exactly in integers.
__global__ void __launch_bounds__(256)
triu_flatten_kernel(const float* __restrict__ M, float* __restrict__ out, int n, long long n_el) {
    const float* Mb = M + (size_t)blockIdx.y * n * n;
    float* ob = out + (size_t)blockIdx.y * n_el;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_el; t += (long long)gridDim.x * blockDim.x) {
        // rows before r hold S(r) = r n - r (r + 1) / 2 elements; largest r with S(r) <= t
        const double nn = 2.0 * n - 1.0;
        int r = (int)((nn - sqrt(nn * nn - 8.0 * (double)t)) * 0.5);
        r = max(0, min(r, n - 2));
        while (r > 0 && (long long)r * n - (long long)r * (r + 1) / 2 > t) --r;
        while ((long long)(r + 1) * n - (long long)(r + 1) * (r + 2) / 2 <= t) ++r;
        const int c = (int)(t - ((long long)r * n - (long long)r * (r + 1) / 2)) + r + 1;
        ob[t] = Mb[(size_t)r * n + c];
    }
}

// ---------------------------------------------------------------- 64 x 64 tile fast paths (4-byte elements)
// The 32 x 32 kernels above move 4 KB per CTA and launch a CTA for every tile of the square (the lower half
// exits at once): launch bound, 27-29 % of HBM at the 268-node batch.  Here a CTA takes one 64 x 64 tile of
// the UPPER tile triangle (15 per 268 x 268 matrix), reads it with float4 row segments, and
//   vectorize      writes, for every column c of the tile, the run of consecutive r (coalesced along r);
//   anti_vectorize reads the rows of the packed upper triangle (contiguous in c), writes tile (I, J) row-wise
//                  and the mirror tile (J, I) from a transposed read of shared memory.
constexpr int kVT = 64, kVP = 65, kVThreads = 256;

__device__ __forceinline__ void upper_tile_of(int idx, int T, int& I, int& J) {
    I = 0;
    while (idx >= T - I) { idx -= T - I; ++I; }
    J = I + idx;
}

__global__ void __launch_bounds__(kVThreads)
vectorize_tiles_kernel(const uint32_t* __restrict__ M, uint32_t* __restrict__ out, int n, long long len, int diag, int T) {
    __shared__ uint32_t tile[kVT * kVP];
    const uint32_t* Mb = M + (size_t)blockIdx.y * n * n;
    uint32_t* ob = out + (size_t)blockIdx.y * len;
    int I, J;
    upper_tile_of(blockIdx.x, T, I, J);
    const int r0 = I * kVT, c0 = J * kVT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int rr = ty + 16 * p, r = r0 + rr, c = c0 + 4 * tx;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < n && c < n) v = __ldg(reinterpret_cast<const uint4*>(Mb + (size_t)r * n + c));      // n % 4 == 0
        tile[rr * kVP + 4 * tx] = v.x; tile[rr * kVP + 4 * tx + 1] = v.y; tile[rr * kVP + 4 * tx + 2] = v.z; tile[rr * kVP + 4 * tx + 3] = v.w;
    }
    __syncthreads();
    // column cc of the tile -> out[start(c) + r0 + rr], rr ascending: lanes walk rr
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;                 // 8 warps
    for (int cc = w; cc < kVT; cc += 8) {
        const int c = c0 + cc;
        if (c >= n) break;
        const long long st = vec_start(c, diag);
        for (int rr = lane; rr < kVT; rr += 32) {
            const int r = r0 + rr;
            if (r < c) ob[st + r] = tile[rr * kVP + cc];
            else if (diag && r == c + 1 && r < n) ob[st + c] = tile[rr * kVP + cc];        // sub-diagonal element (c+1, c)
        }
    }
    // the sub-diagonal element of the last column of a diagonal tile lives in the tile below it
    if (diag && I == J && threadIdx.x == 0) {
        const int c = c0 + kVT - 1, r = c + 1;
        if (r < n) ob[vec_start(c, 1) + c] = Mb[(size_t)r * n + c];
    }
}

__global__ void __launch_bounds__(kVThreads)
anti_vectorize_tiles_kernel(const float* __restrict__ vec, float* __restrict__ out, int n, long long vec_len, int diag, int T) {
    __shared__ float tile[kVT * kVP];
    const float* vb = vec + (size_t)blockIdx.y * vec_len;
    float* ob = out + (size_t)blockIdx.y * n * n;
    const long long n_el = (long long)n * (n - 1) / 2;
    int I, J;
    upper_tile_of(blockIdx.x, T, I, J);
    const int r0 = I * kVT, c0 = J * kVT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // value of the symmetric matrix at (r, c) of tile (I, J): from the packed row r for r < c, from row c for r > c
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int rr = ty + 16 * p, r = r0 + rr;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < n) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + 4 * tx + u;
                if (c < n) {
                    if (r < c) o[u] = __ldg(vb + (long long)r * n - (long long)r * (r + 1) / 2 + (c - r - 1));
                    else if (r > c) o[u] = __ldg(vb + (long long)c * n - (long long)c * (c + 1) / 2 + (r - c - 1));   // diagonal tile only
                    else if (diag) o[u] = __ldg(vb + n_el + r);
                }
            }
            if (c0 + 4 * tx < n) *reinterpret_cast<float4*>(ob + (size_t)r * n + c0 + 4 * tx) = make_float4(o[0], o[1], o[2], o[3]);
        }
        tile[rr * kVP + 4 * tx] = o[0]; tile[rr * kVP + 4 * tx + 1] = o[1]; tile[rr * kVP + 4 * tx + 2] = o[2]; tile[rr * kVP + 4 * tx + 3] = o[3];
    }
    if (I == J) return;
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {                       // mirror tile (J, I): out[c][r] = value(r, c)
        const int cc = ty + 16 * p, c = c0 + cc;
        if (c < n && r0 + 4 * tx < n) {
            *reinterpret_cast<float4*>(ob + (size_t)c * n + r0 + 4 * tx) =
                make_float4(tile[(4 * tx) * kVP + cc], tile[(4 * tx + 1) * kVP + cc], tile[(4 * tx + 2) * kVP + cc], tile[(4 * tx + 3) * kVP + cc]);
        }
    }
}

template <typename T>
static int launch_vectorize(const void* M, void* out, int batch, int n, int diag, cudaStream_t st) {
    const int tiles = (n + 31) / 32;
    dim3 grid(tiles, tiles, batch);
    vectorize_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(M), static_cast<T*>(out), n,
                                              gsr_vectorize_len(n, diag), diag);
    return launched("vectorize");
}

}  // namespace gsr

using namespace gsr;

extern "C" long long gsr_vectorize_len(int n, int include_diagonal) {
    if (n <= 0) return 0;
    return (long long)n * (n - 1) / 2 + (include_diagonal ? (n - 1) : 0);
}

extern "C" int gsr_vectorize(const void* matrix, void* out, int batch, int n, int elem_size,
                             int include_diagonal, void* stream) {
    if (!matrix || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "vectorize: bad argument");
    if (batch == 0 || n == 1) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "vectorize: batch > 65535");
    cudaStream_t st = (cudaStream_t)stream;
    const int d = include_diagonal ? 1 : 0;
    if (elem_size == 4 && n % 4 == 0 && n >= 2 * kVT && aligned16(matrix)) {
        const int T = (n + kVT - 1) / kVT;
        vectorize_tiles_kernel<<<dim3(T * (T + 1) / 2, batch), kVThreads, 0, st>>>(
            static_cast<const uint32_t*>(matrix), static_cast<uint32_t*>(out), n, gsr_vectorize_len(n, d), d, T);
        return launched("vectorize");
    }
    switch (elem_size) {
        case 1: return launch_vectorize<uint8_t>(matrix, out, batch, n, d, st);
        case 2: return launch_vectorize<uint16_t>(matrix, out, batch, n, d, st);
        case 4: return launch_vectorize<uint32_t>(matrix, out, batch, n, d, st);
        case 8: return launch_vectorize<uint64_t>(matrix, out, batch, n, d, st);
        case 16: return launch_vectorize<uint4>(matrix, out, batch, n, d, st);
        default: return fail(GSR_E_UNSUPPORTED, "vectorize: elem_size must be 1, 2, 4, 8 or 16");
    }
}

extern "C" int gsr_anti_vectorize(const float* vec, float* out, int batch, int n, long long vec_len,
                                  int include_diagonal, void* stream) {
    if (!vec || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "anti_vectorize: bad argument");
    const long long n_el = (long long)n * (n - 1) / 2;
    if (vec_len < n_el) return fail(GSR_E_ARG, "anti_vectorize: vector shorter than n(n-1)/2");
    if (include_diagonal && vec_len < n_el + n) return fail(GSR_E_ARG, "anti_vectorize: vector too short for the diagonal");
    if (batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "anti_vectorize: batch > 65535");
    if (n % 4 == 0 && n >= 2 * kVT && aligned16(out)) {
        const int T = (n + kVT - 1) / kVT;
        anti_vectorize_tiles_kernel<<<dim3(T * (T + 1) / 2, batch), kVThreads, 0, (cudaStream_t)stream>>>(vec, out, n, vec_len, include_diagonal ? 1 : 0, T);
        return launched("anti_vectorize");
    }
    const long long total = (long long)n * n;
    const int bx = (int)min((long long)64, (total + 255) / 256);
    anti_vectorize_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(vec, out, n, vec_len, include_diagonal ? 1 : 0);
    return launched("anti_vectorize");
}

extern "C" int gsr_triu_flatten(const float* matrix, float* out, int batch, int n, void* stream) {
    if (!matrix || !out || batch < 0 || n <= 0) return fail(GSR_E_ARG, "triu_flatten: bad argument");
    if (batch == 0 || n == 1) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "triu_flatten: batch > 65535");
    const long long n_el = (long long)n * (n - 1) / 2;
    const int bx = (int)min((long long)64, (n_el + 255) / 256);
    triu_flatten_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(matrix, out, n, n_el);
    return launched("triu_flatten");
}
