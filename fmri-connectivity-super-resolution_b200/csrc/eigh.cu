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
//   * polish (when a second n x LD buffer fits in shared memory, n <= ~170): the
//     shift puts sigma-sized entries on the diagonal of G, so ~1500 fp32 rotations
//     per column leave eigenvector errors of O(eps*sigma*sqrt(1500)/gap) ~ 2e-3 in
//     the clustered bulk -- 50x LAPACK's.  The second phase recomputes H = A.U from
//     the input and runs the implicit two-sided Jacobi (pivots u_p.h_q, rotations
//     applied to the columns of U and H alike) for the few sweeps it takes; its
//     rounding is relative to the small entries of A.U, which brings the basis to
//     LAPACK's accuracy class (measured: U 2.5e-5..7e-5 vs 1e-5..3.6e-5 abs).
//   * eigenvalues are Rayleigh quotients u^T A u (from H after the polish, from the
//     UPPER triangle of the input otherwise); column norms are never used for them
//     (they drift by O(1e-5) over the rotations).
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

// Rotation (c, s) that zeroes the pivot of [[aa, ab], [ab, bb]].  MUFU-based (rsqrtf,
// __fdividef): a (c, s) pair that is off by a common factor 1 + O(ulp) scales both rotated
// columns alike, which one-sided Jacobi does not care about (directions and spans are kept;
// columns are normalised at the end), so IEEE-exact roots / divisions would only lengthen
// the dependent chain that every round waits on.
__device__ __forceinline__ void jacobi_cs(float aa, float bb, float ab, float& c, float& s, float& t) {
    const float zeta = __fdividef(bb - aa, 2.0f * ab);
    t = __fdividef(copysignf(1.0f, zeta), fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
    c = rsqrtf(fmaf(t, t, 1.0f));
    s = c * t;
}

// KR     : rows per lane / 4 (column pitch LD = 32 KR + 4 floats)
// POLISH : second phase (needs a second n x LD buffer in shared memory), see header comment
template <int KR, bool POLISH>
__global__ void __launch_bounds__(128 * KR, 1)
eigh_jacobi_kernel(const float* __restrict__ Ain, float* __restrict__ Wout, float* __restrict__ Uout,
                   int32_t* __restrict__ info, int n, int canonical, float tol, int max_sweeps) {
    constexpr int LD = 32 * KR + 4;
    constexpr int R = 4 * KR;
    extern __shared__ __align__(16) float smem[];
    float* G = smem;                                            // [n][LD], column c at G + c*LD
    float* H = G + (size_t)n * LD;                              // [n][LD] (POLISH only)
    float* lam = POLISH ? H + (size_t)n * LD : H;               // [n]
    float* sgn = lam + n;                                       // [n]
    float* nrm = sgn + n;                                       // [n]  cached squared column norms
    int* inv = reinterpret_cast<int*>(nrm + n);                 // [n]
    __shared__ float s_red[32];
    __shared__ float s_sigma;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = tid >> 3, j = tid & 7;                        // pair slot, lane inside the slot
    const unsigned gmask = 0xFFu << (lane & 24);                // the 8 lanes of this slot
    const int n_pad = (n + 1) & ~1, npairs = n_pad >> 1, m = n_pad - 1;
    const float* A = Ain + (size_t)blockIdx.x * n * n;
    float* Wg = Wout + (size_t)blockIdx.x * n;
    float* Ug = Uout + (size_t)blockIdx.x * n * n;

    // ---- load: zero the padding, mirror the upper triangle, detect non-finite input.  The
    //      symmetrised matrix is also parked in the U output buffer (scratch until the final
    //      store) so that later passes can read its rows contiguously.
    for (int e = tid; e < n * LD; e += nthr) G[e] = 0.f;
    __syncthreads();
    int bad = 0;
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, c = e - i * n;
        if (c >= i) {
            const float v = __ldg(A + e);
            if (!isfinite(v)) bad = 1;
            G[(size_t)i * LD + c] = v;                          // column i, row c  (= A_sym[c][i])
            Ug[e] = v;
            if (c > i) { G[(size_t)c * LD + i] = v; Ug[(size_t)c * n + i] = v; }
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

    // ---- phase 1: cyclic one-sided Jacobi sweeps on G
    int sweeps = 0, converged = 0;
    while (sweeps < max_sweeps) {
        // exact squared norms once per sweep; inside the sweep they follow aa -= t ab, bb += t ab
        for (int pass = 0; pass < 2; ++pass) {
            const int k = 2 * g + pass;
            if (g < npairs && k < n) {
                float u[R];
                load_col<KR>(G + (size_t)k * LD, j, u);
                float ss = 0.f;
#pragma unroll
                for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
                ss = group_sum8(ss, gmask);
                if (j == 0) nrm[k] = ss;
            }
        }
        __syncthreads();
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
                const float aa = nrm[p], bb = nrm[q];
                float ab0 = 0.f, ab1 = 0.f;
#pragma unroll
                for (int e = 0; e < R; e += 2) {
                    ab0 = fmaf(a[e], b[e], ab0);
                    ab1 = fmaf(a[e + 1], b[e + 1], ab1);
                }
                const float ab = group_sum8(ab0 + ab1, gmask);
                if (ab * ab > tol * tol * aa * bb) {
                    float c, s, t;
                    jacobi_cs(aa, bb, ab, c, s, t);
#pragma unroll
                    for (int e = 0; e < R; ++e) {
                        const float x = a[e], y = b[e];
                        a[e] = c * x - s * y;
                        b[e] = s * x + c * y;
                    }
                    store_col<KR>(cp, j, a);
                    store_col<KR>(cq, j, b);
                    if (j == 0) { nrm[p] = aa - t * ab; nrm[q] = bb + t * ab; }
                    rotated = 1;
                }
            }
            __syncthreads();
        }
        ++sweeps;
        if (!__syncthreads_or(rotated)) { converged = 1; break; }
    }

    // ---- normalise the columns: U
    for (int pass = 0; pass < 2; ++pass) {
        const int k = 2 * g + pass;
        if (g < npairs && k < n) {
            float u[R];
            float* ck = G + (size_t)k * LD;
            load_col<KR>(ck, j, u);
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
            ss = group_sum8(ss, gmask);
            const float inv_norm = 1.0f / sqrtf(ss);
#pragma unroll
            for (int e = 0; e < R; ++e) u[e] *= inv_norm;
            store_col<KR>(ck, j, u);
        }
    }
    __syncthreads();

    const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ug) & 15u) == 0);
    if (POLISH) {
        // ---- phase 2a: H = A_sym U straight from the input (fresh, no accumulated rounding);
        //      slot g produces columns 2g and 2g+1, lane j the rows it owns.
        {
            const int k0 = 2 * g, k1 = 2 * g + 1;
            const bool act0 = g < npairs && k0 < n, act1 = g < npairs && k1 < n;
            if (act0) {
                float h0[R], h1[R];
#pragma unroll
                for (int e = 0; e < R; ++e) { h0[e] = 0.f; h1[e] = 0.f; }
                const float* c0 = G + (size_t)k0 * LD;
                const float* c1 = G + (size_t)(act1 ? k1 : k0) * LD;
                for (int mm = 0; mm < n; ++mm) {
                    const float u0 = c0[mm], u1 = c1[mm];
                    const float* Ar = Ug + (size_t)mm * n;
#pragma unroll
                    for (int kk = 0; kk < KR; ++kk) {
                        const int i0 = 4 * j + 32 * kk;
                        float4 av;
                        if (vec_ok && i0 + 3 < n) av = *reinterpret_cast<const float4*>(Ar + i0);
                        else {
                            av.x = (i0 + 0 < n) ? Ar[i0 + 0] : 0.f;
                            av.y = (i0 + 1 < n) ? Ar[i0 + 1] : 0.f;
                            av.z = (i0 + 2 < n) ? Ar[i0 + 2] : 0.f;
                            av.w = (i0 + 3 < n) ? Ar[i0 + 3] : 0.f;
                        }
                        h0[4 * kk + 0] = fmaf(av.x, u0, h0[4 * kk + 0]); h1[4 * kk + 0] = fmaf(av.x, u1, h1[4 * kk + 0]);
                        h0[4 * kk + 1] = fmaf(av.y, u0, h0[4 * kk + 1]); h1[4 * kk + 1] = fmaf(av.y, u1, h1[4 * kk + 1]);
                        h0[4 * kk + 2] = fmaf(av.z, u0, h0[4 * kk + 2]); h1[4 * kk + 2] = fmaf(av.z, u1, h1[4 * kk + 2]);
                        h0[4 * kk + 3] = fmaf(av.w, u0, h0[4 * kk + 3]); h1[4 * kk + 3] = fmaf(av.w, u1, h1[4 * kk + 3]);
                    }
                }
                store_col<KR>(H + (size_t)k0 * LD, j, h0);
                if (act1) store_col<KR>(H + (size_t)k1 * LD, j, h1);
            }
        }
        __syncthreads();
        // ---- phase 2b: implicit two-sided Jacobi on T = U^T H (pivots are dot products of a
        //      U column with an H column; the rotation is applied to the columns of both).
        const float theta_tol = 2e-6f, noise = 3e-8f;
        for (int psweep = 0; psweep < 8; ++psweep) {
            int rotated = 0;
            for (int r = 0; r < m; ++r) {
                int p, q;
                if (g == 0) { p = n_pad - 1; q = r; }
                else { p = r + g; if (p >= m) p -= m; q = r - g; if (q < 0) q += m; }
                if (g < npairs && p < n && q < n) {
                    float up[R], uq[R], hx[R];
                    load_col<KR>(G + (size_t)p * LD, j, up);
                    load_col<KR>(G + (size_t)q * LD, j, uq);
                    load_col<KR>(H + (size_t)p * LD, j, hx);            // h_p
                    float tpp = 0.f, tqp = 0.f, tqq = 0.f, tpq = 0.f;
#pragma unroll
                    for (int e = 0; e < R; ++e) { tpp = fmaf(up[e], hx[e], tpp); tqp = fmaf(uq[e], hx[e], tqp); }
                    load_col<KR>(H + (size_t)q * LD, j, hx);            // h_q
#pragma unroll
                    for (int e = 0; e < R; ++e) { tqq = fmaf(uq[e], hx[e], tqq); tpq = fmaf(up[e], hx[e], tpq); }
                    tpp = group_sum8(tpp, gmask); tqq = group_sum8(tqq, gmask);
                    const float off = 0.5f * group_sum8(tpq + tqp, gmask);
                    const float thr = fmaxf(theta_tol * fabsf(tqq - tpp), noise * (fabsf(tpp) + fabsf(tqq)));
                    if (fabsf(off) > thr) {
                        float c, s, t;
                        jacobi_cs(tpp, tqq, off, c, s, t);
#pragma unroll
                        for (int e = 0; e < R; ++e) {
                            const float x = up[e], y = uq[e];
                            up[e] = c * x - s * y;
                            uq[e] = s * x + c * y;
                        }
                        store_col<KR>(G + (size_t)p * LD, j, up);
                        store_col<KR>(G + (size_t)q * LD, j, uq);
                        load_col<KR>(H + (size_t)p * LD, j, up);        // reuse registers: h_p
#pragma unroll
                        for (int e = 0; e < R; ++e) {
                            const float x = up[e], y = hx[e];
                            up[e] = c * x - s * y;
                            hx[e] = s * x + c * y;
                        }
                        store_col<KR>(H + (size_t)p * LD, j, up);
                        store_col<KR>(H + (size_t)q * LD, j, hx);
                        rotated = 1;
                    }
                }
                __syncthreads();
            }
            if (!__syncthreads_or(rotated)) break;
        }
    }

    // ---- eigenvalues (Rayleigh quotients), unit columns, sign pivots
    for (int pass = 0; pass < 2; ++pass) {
        const int k = 2 * g + pass;
        const bool active = (g < npairs) && (k < n);
        float u[R];
        float acc = 0.f, uu = 1.f;
        if (active) {
            float* ck = G + (size_t)k * LD;
            load_col<KR>(ck, j, u);
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
            ss = group_sum8(ss, gmask);
            uu = ss;
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
            if (POLISH) {
                float hk[R];
                load_col<KR>(H + (size_t)k * LD, j, hk);
#pragma unroll
                for (int e = 0; e < R; ++e) acc = fmaf(u[e], hk[e], acc);       // u normalised, h = A u_old
                acc *= inv_norm;                                                 // -> u^T A u
            }
        }
        if (!POLISH) {
            __syncwarp();
            // lambda_k = sum_i u_i (A_ii u_i + 2 sum_{mm > i} A[i][mm] u_mm) from the upper triangle
            if (active) {
                const float* ck = G + (size_t)k * LD;
                for (int i = 0; i < n; ++i) {
                    const float ui = ck[i];
                    const float* Ar = A + (size_t)i * n;
                    float part = 0.f;
#pragma unroll
                    for (int kk = 0; kk < KR; ++kk) {
                        const int m0 = 4 * j + 32 * kk;
                        if (m0 + 3 < i) continue;
                        float4 av;
                        av.x = (m0 + 0 < n) ? __ldg(Ar + m0 + 0) : 0.f;
                        av.y = (m0 + 1 < n) ? __ldg(Ar + m0 + 1) : 0.f;
                        av.z = (m0 + 2 < n) ? __ldg(Ar + m0 + 2) : 0.f;
                        av.w = (m0 + 3 < n) ? __ldg(Ar + m0 + 3) : 0.f;
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
        }
        if (active) {
            acc = group_sum8(acc, gmask);
            if (j == 0) lam[k] = acc;
        }
        (void)uu;
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
    const size_t small = 4 * (size_t)n * sizeof(float);
    const size_t one = (size_t)n * LD * sizeof(float) + small, two = 2 * (size_t)n * LD * sizeof(float) + small;
    if (threads > 1024 || one > 227 * 1024) return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: n too large for one CTA");
    const bool polish = two <= 227 * 1024;
    const size_t smem = polish ? two : one;
    cudaError_t e = polish
        ? cudaFuncSetAttribute(eigh_jacobi_kernel<KR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(eigh_jacobi_kernel<KR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "eigh_jacobi: cudaFuncSetAttribute failed");
    if (polish) eigh_jacobi_kernel<KR, true><<<batch, threads, smem, st>>>(A, w, U, info, n, canonical, tol, max_sweeps);
    else        eigh_jacobi_kernel<KR, false><<<batch, threads, smem, st>>>(A, w, U, info, n, canonical, tol, max_sweeps);
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
