// Batched symmetric eigendecomposition for the GSR layer (models/layers.py:27) and
// the HR-target basis (utils/gsr_train.py:34): torch.linalg.eigh(A, UPLO='U').
//
// One-sided (Hestenes) Jacobi, one CTA per graph, the whole working matrix
// resident in shared memory for the life of the kernel:
//   * A is indefinite (normalised adjacency: lambda_max = 1, bulk around 0), so
//     the iteration runs on G = A + sigma*I with sigma = 1.0625 * (Gershgorin
//     bound), which is positive definite.  For a symmetric PD matrix the column
//     rotations that orthogonalise G converge to G.V = U.diag(lambda + sigma)
//     with V -> U, so no separate eigenvector accumulator is kept: the
//     normalised columns ARE the eigenvectors.
//   * columns are stored contiguously (ld = 32*KR + 4 floats); 8 lanes own one
//     column pair, 4*KR rows each, read and written as conflict-free float4s;
//     the three dot products (a.a, b.b, a.b) are reduced with 3 xor-shuffles.
//   * pairs follow the round-robin ("circle") tournament, computed analytically
//     from the round number: n_pad/2 disjoint pairs rotate concurrently, n_pad-1
//     rounds per sweep, one __syncthreads per round.
//   * eigenvalues are Rayleigh quotients u^T A u taken from the UPPER triangle
//     of the input (column norms drift by O(1e-5) over ~1500 fp32 rotations; the
//     quotient is accurate to O(eps * |A|)).
//   * epilogue: rank by counting (ascending), optional sign canonicalisation
//     (largest-|.| component positive, lowest row on ties), coalesced store.
#include "common.cuh"

namespace gsr {

constexpr int kGroup = 8;        // lanes per column pair

// Butterfly over the 8 lanes of one pair slot.  `mask` names only those 8 lanes, so
// slots of the same warp may sit in different branches.  Every lane ends with the
// bitwise-identical sum (each level adds the same two operands in both partners).
__device__ __forceinline__ float group_sum8(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
}

template <int KR>
__device__ __forceinline__ void load_col(const float* __restrict__ col, int j, float (&v)[4 * KR]) {
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const float4 t = *reinterpret_cast<const float4*>(col + 4 * j + 32 * k);
        v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
}
template <int KR>
__device__ __forceinline__ void store_col(float* __restrict__ col, int j, const float (&v)[4 * KR]) {
#pragma unroll
    for (int k = 0; k < KR; ++k)
        *reinterpret_cast<float4*>(col + 4 * j + 32 * k) =
            make_float4(v[4 * k + 0], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

template <int KR>
__global__ void __launch_bounds__(128 * KR, 1)
eigh_jacobi_kernel(const float* __restrict__ Ain, float* __restrict__ Wout, float* __restrict__ Uout,
                   int32_t* __restrict__ info, int n, int canonical, float tol, int max_sweeps) {
    constexpr int LD = 32 * KR + 4;
    constexpr int R = 4 * KR;
    extern __shared__ __align__(16) float smem[];
    float* G = smem;                                  // [n][LD], column c at G + c*LD
    float* lam = G + (size_t)n * LD;                  // [n]
    float* sgn = lam + n;                             // [n]
    int* rank = reinterpret_cast<int*>(sgn + n);      // [n]
    int* inv = rank + n;                              // [n]
    __shared__ float s_red[32];
    __shared__ float s_sigma;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = tid >> 3, j = tid & 7;              // pair slot, lane inside the slot
    const unsigned gmask = 0xFFu << (lane & 24);      // the 8 lanes of this slot
    const int n_pad = (n + 1) & ~1, npairs = n_pad >> 1, m = n_pad - 1;
    const float* A = Ain + (size_t)blockIdx.x * n * n;
    float* Wg = Wout + (size_t)blockIdx.x * n;
    float* Ug = Uout + (size_t)blockIdx.x * n * n;

    // ---- load: zero the padded columns, mirror the upper triangle, find non-finite input
    for (int e = tid; e < n * LD; e += nthr) G[e] = 0.f;
    __syncthreads();
    int bad = 0;
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, c = e - i * n;
        if (c >= i) {
            const float v = __ldg(A + e);
            if (!isfinite(v)) bad = 1;
            G[(size_t)i * LD + c] = v;                // column i, row c  (= A_sym[c][i])
            if (c > i) G[(size_t)c * LD + i] = v;     // column c, row i
        }
    }
    if (__syncthreads_or(bad)) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int e = tid; e < n * n; e += nthr) Ug[e] = qnan;
        for (int e = tid; e < n; e += nthr) Wg[e] = qnan;
        if (tid == 0) info[blockIdx.x] = -1;
        return;
    }
    // ---- sigma = 1.0625 * max_i sum_j |A_sym[i][j]|  (Gershgorin), G = A + sigma I
    float rowmax = 0.f;
    for (int c = warp; c < n; c += (nthr >> 5)) {
        float acc = 0.f;
        for (int i = lane; i < n; i += 32) acc += fabsf(G[(size_t)c * LD + i]);
        acc = warp_sum(acc);
        rowmax = fmaxf(rowmax, acc);
    }
    if (lane == 0) s_red[warp] = rowmax;
    __syncthreads();
    if (tid == 0) {
        float mx = 0.f;
        for (int w2 = 0; w2 < (nthr >> 5); ++w2) mx = fmaxf(mx, s_red[w2]);
        s_sigma = (mx > 0.f) ? 1.0625f * mx : 1.0f;
    }
    __syncthreads();
    const float sigma = s_sigma;
    for (int c = tid; c < n; c += nthr) G[(size_t)c * LD + c] += sigma;
    __syncthreads();

    // ---- cyclic one-sided Jacobi sweeps
    int sweeps = 0, converged = 0;
    while (sweeps < max_sweeps) {
        int rotated = 0;
        for (int r = 0; r < m; ++r) {
            int p, q;
            if (g == 0) { p = n_pad - 1; q = r; }
            else { p = r + g; if (p >= m) p -= m; q = r - g; if (q < 0) q += m; }
            if (g < npairs && p < n && q < n) {
                float a[R], b[R];
                float* cp = G + (size_t)p * LD;
                float* cq = G + (size_t)q * LD;
                load_col<KR>(cp, j, a);
                load_col<KR>(cq, j, b);
                float aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
                for (int e = 0; e < R; ++e) {
                    aa = fmaf(a[e], a[e], aa);
                    bb = fmaf(b[e], b[e], bb);
                    ab = fmaf(a[e], b[e], ab);
                }
                aa = group_sum8(aa, gmask); bb = group_sum8(bb, gmask); ab = group_sum8(ab, gmask);
                if (fabsf(ab) > tol * sqrtf(aa * bb)) {
                    const float zeta = (bb - aa) / (2.0f * ab);
                    const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
                    const float c = 1.0f / sqrtf(1.0f + t * t);
                    const float s = c * t;
#pragma unroll
                    for (int e = 0; e < R; ++e) {
                        const float x = a[e], y = b[e];
                        a[e] = c * x - s * y;
                        b[e] = s * x + c * y;
                    }
                    store_col<KR>(cp, j, a);
                    store_col<KR>(cq, j, b);
                    rotated = 1;
                }
            }
            __syncthreads();
        }
        ++sweeps;
        if (!__syncthreads_or(rotated)) { converged = 1; break; }
    }

    // ---- normalise columns (U), sign pivot, Rayleigh quotient against the upper triangle
    const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
    for (int pass = 0; pass < 2; ++pass) {
        const int k = 2 * g + pass;
        const bool active = (g < npairs) && (k < n);
        float u[R];
        if (active) {
            float* ck = G + (size_t)k * LD;
            load_col<KR>(ck, j, u);
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
            ss = group_sum8(ss, gmask);
            const float inv_norm = 1.0f / sqrtf(ss);
            float best = -1.f; int best_i = 0; float best_v = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) {
                u[e] *= inv_norm;
                const int row = 4 * j + 32 * (e >> 2) + (e & 3);
                const float mag = fabsf(u[e]);
                if (row < n && mag > best) { best = mag; best_i = row; best_v = u[e]; }
            }
            store_col<KR>(ck, j, u);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(gmask, best, o);
                const int oi = __shfl_xor_sync(gmask, best_i, o);
                const float ov = __shfl_xor_sync(gmask, best_v, o);
                if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_v = ov; }
            }
            if (j == 0) sgn[k] = (canonical && best_v < 0.f) ? -1.0f : 1.0f;
        }
        __syncwarp();
        // lambda_k = sum_i u_i (A_ii u_i + 2 sum_{mm > i} A[i][mm] u_mm); lane owns mm in its rows
        float acc = 0.f;
        if (active) {
            const float* ck = G + (size_t)k * LD;
            for (int i = 0; i < n; ++i) {
                const float ui = ck[i];               // own column: written by this group only
                const float* Ar = A + (size_t)i * n;
                float part = 0.f;
#pragma unroll
                for (int kk = 0; kk < KR; ++kk) {
                    const int m0 = 4 * j + 32 * kk;
                    if (m0 + 3 < i) continue;         // whole chunk below the diagonal
                    float4 av;
                    if (vec_ok && m0 + 3 < n) av = __ldg(reinterpret_cast<const float4*>(Ar + m0));
                    else {
                        av.x = (m0 + 0 < n) ? __ldg(Ar + m0 + 0) : 0.f;
                        av.y = (m0 + 1 < n) ? __ldg(Ar + m0 + 1) : 0.f;
                        av.z = (m0 + 2 < n) ? __ldg(Ar + m0 + 2) : 0.f;
                        av.w = (m0 + 3 < n) ? __ldg(Ar + m0 + 3) : 0.f;
                    }
                    const float f0 = (m0 + 0 > i) ? 2.f : ((m0 + 0 == i) ? 1.f : 0.f);
                    const float f1 = (m0 + 1 > i) ? 2.f : ((m0 + 1 == i) ? 1.f : 0.f);
                    const float f2 = (m0 + 2 > i) ? 2.f : ((m0 + 2 == i) ? 1.f : 0.f);
                    const float f3 = (m0 + 3 > i) ? 2.f : ((m0 + 3 == i) ? 1.f : 0.f);
                    part = fmaf(av.x * f0, u[4 * kk + 0], part);
                    part = fmaf(av.y * f1, u[4 * kk + 1], part);
                    part = fmaf(av.z * f2, u[4 * kk + 2], part);
                    part = fmaf(av.w * f3, u[4 * kk + 3], part);
                }
                acc = fmaf(ui, part, acc);
            }
        }
        acc = group_sum8(acc, gmask);
        if (active && j == 0) lam[k] = acc;
    }
    __syncthreads();

    // ---- ascending order by counting; ties keep column order
    for (int k = tid; k < n; k += nthr) {
        const float lk = lam[k];
        int rk = 0;
        for (int q2 = 0; q2 < n; ++q2) {
            const float lq = lam[q2];
            rk += (lq < lk) || (lq == lk && q2 < k);
        }
        rank[k] = rk;
        inv[rk] = k;
        Wg[rk] = lk;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, c = e - i * n;
        const int k = inv[c];
        Ug[e] = G[(size_t)k * LD + i] * sgn[k];
    }
    if (tid == 0) info[blockIdx.x] = converged ? sweeps : -2;
}

template <int KR>
static int launch_eigh(const float* A, float* w, float* U, int32_t* info, int batch, int n,
                       int canonical, float tol, int max_sweeps, cudaStream_t st) {
    constexpr int LD = 32 * KR + 4;
    const int n_pad = (n + 1) & ~1, npairs = n_pad / 2;
    const int threads = 32 * ((npairs * kGroup + 31) / 32);
    const size_t smem = ((size_t)n * LD + 4 * (size_t)n) * sizeof(float);
    if (threads > 1024 || smem > 227 * 1024) return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: n too large for one CTA");
    cudaError_t e = cudaFuncSetAttribute(eigh_jacobi_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "eigh_jacobi: cudaFuncSetAttribute failed");
    eigh_jacobi_kernel<KR><<<batch, threads, smem, st>>>(A, w, U, info, n, canonical, tol, max_sweeps);
    return launched("eigh_jacobi");
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_eigh_jacobi_f32(const float* A, float* w, float* U, int32_t* info,
                                   int batch, int n, int canonical_sign, float tol, int max_sweeps,
                                   void* stream) {
    if (!A || !w || !U || !info || batch < 0 || n <= 0) return fail(GSR_E_ARG, "eigh_jacobi: bad argument");
    if (batch == 0) return 0;
    if (tol <= 0.f) tol = 1e-6f;
    if (max_sweeps <= 0) max_sweeps = 30;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > GSR_EIGH_MAX_N_CTA) return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: n > 224 needs the cluster kernel (not built yet)");
    switch ((n + 31) / 32) {
        case 1: return launch_eigh<1>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 2: return launch_eigh<2>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 3: return launch_eigh<3>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 4: return launch_eigh<4>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 5: return launch_eigh<5>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 6: return launch_eigh<6>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        case 7: return launch_eigh<7>(A, w, U, info, batch, n, canonical_sign, tol, max_sweeps, st);
        default: return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: unsupported n");
    }
}
