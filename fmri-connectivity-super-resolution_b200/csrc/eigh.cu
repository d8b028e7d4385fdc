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
//   * columns are stored contiguously (ld = 32*KR + 4 floats); 8 (or 16) lanes own one
//     column pair, 4*KR (2*KR) rows each, read and written as conflict-free float4s (float2s);
//     the pivot dot product is reduced with 3 xor-shuffles; squared column norms
//     are cached (exact once per sweep, updated as aa -= t*ab, bb += t*ab).
//   * pair ordering: odd-even transposition.  Columns sit at positions 0..n-1; even steps
//     pair positions (2g, 2g+1), odd steps (2g+1, 2g+2), and the two columns swap positions
//     after every rotation, so n_pad steps bring every pair together exactly once.  Slot g
//     keeps the column at position 2g+1 in REGISTERS for the whole sweep and exchanges only
//     the other one through shared memory: half the LDS / STS instructions and bytes of the
//     round-robin ("circle") ordering at the same occupancy -- the kernel is bound by the
//     LSU / shared-memory pipe (ncu: mio_throttle + short_scoreboard, data pipe 66 % busy),
//     and this took 1024 graphs from 19.6 to 14.6 ms.  (Tried and measured slower on B200:
//     2-column block rotations with 8 or 16 lanes per slot, and neighbour-flag wavefront
//     synchronisation instead of the per-step barrier.)  The polish phase and the
//     global-workspace variant use the round-robin ordering, n_pad-1 rounds per sweep,
//     computed analytically from the round number.
//   * polish (when a second n x LD buffer fits in shared memory, n <= ~170): the
//     shift puts sigma-sized entries on the diagonal of G, so ~1500 fp32 rotations
//     per column leave eigenvector errors of O(eps*sigma*sqrt(1500)/gap) ~ 2e-3 in
//     the clustered bulk -- 50x LAPACK's.  The second phase recomputes H = A.U from
//     the input and runs the implicit two-sided Jacobi (pivots u_p.h_q, rotations
//     applied to the columns of U and H alike) for the one or two sweeps it takes;
//     its rounding is relative to the small entries of A.U, which brings the basis
//     to LAPACK's accuracy class (measured on B200: max |dU| 2e-5..9e-5 against a
//     float64 decomposition, LAPACK ssyevd: 1e-5..3e-4 on the same matrices).
//   * eigenvalues are Rayleigh quotients u^T A u (from H after the polish, from the
//     UPPER triangle of the input otherwise); column norms are never used for them
//     (they drift by O(1e-5) over the rotations).
//   * epilogue: rank by counting (ascending), optional sign canonicalisation
//     (largest-|.| component positive, lowest row on ties), coalesced store.
//   * n > 224 (the 268-node HR target): the working matrix does not fit in one
//     CTA's shared memory; the same kernel then keeps it in a caller-supplied
//     global workspace (L2 resident, .cg accesses), pair slots are looped over.
//     This is the hoisted, once-per-dataset path; the cluster /
//     distributed-shared-memory variant is future work.
#include "common.cuh"

namespace gsr {

constexpr int kGroup = 8;        // lanes per column pair

// Butterfly over the 8 lanes of one pair slot.  `mask` names only those 8 lanes, so
// slots of the same warp may sit in different branches.  Every lane ends with the
// bitwise-identical sum (each level adds the same two operands in both partners).
template <int LG>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
    if (LG == 16) v += __shfl_xor_sync(mask, v, 8);
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
}

// GLOBAL: the working matrix lives in a global-memory workspace; .cg accesses (L2) keep the
// CTA coherent across its own barriers.
// LG lanes share a column: lane j owns rows VEC*j + 32*k + (0..VEC-1), VEC = 32 / LG, so that the
// LG lanes of one access cover 128 contiguous bytes (conflict free).
template <int KR, bool GLOBAL, int LG>
__device__ __forceinline__ void load_col(const float* col, int j, float (&v)[(32 / LG) * KR]) {
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        if (LG == 8) {
            const float4* ptr = reinterpret_cast<const float4*>(col + 4 * j + 32 * k);
            const float4 t = GLOBAL ? __ldcg(ptr) : *ptr;
            v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        } else {
            const float2* ptr = reinterpret_cast<const float2*>(col + 2 * j + 32 * k);
            const float2 t = GLOBAL ? __ldcg(ptr) : *ptr;
            v[2 * k + 0] = t.x; v[2 * k + 1] = t.y;
        }
    }
}
template <int KR, bool GLOBAL, int LG>
__device__ __forceinline__ void store_col(float* col, int j, const float (&v)[(32 / LG) * KR]) {
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        if (LG == 8) {
            float4* ptr = reinterpret_cast<float4*>(col + 4 * j + 32 * k);
            const float4 t = make_float4(v[4 * k + 0], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            if (GLOBAL) __stcg(ptr, t); else *ptr = t;
        } else {
            float2* ptr = reinterpret_cast<float2*>(col + 2 * j + 32 * k);
            const float2 t = make_float2(v[2 * k + 0], v[2 * k + 1]);
            if (GLOBAL) __stcg(ptr, t); else *ptr = t;
        }
    }
}
template <bool GLOBAL>
__device__ __forceinline__ float load_1(const float* p) { return GLOBAL ? __ldcg(p) : *p; }
template <bool GLOBAL>
__device__ __forceinline__ void store_1(float* p, float v) { if (GLOBAL) __stcg(p, v); else *p = v; }

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

// Round-robin pairing: slot 0 pairs the fixed player n_pad-1 with r; slot i pairs r+i with r-i.
__device__ __forceinline__ void rr_pair(int slot, int r, int m, int n_pad, int& p, int& q) {
    if (slot == 0) { p = n_pad - 1; q = r; }
    else { p = r + slot; if (p >= m) p -= m; q = r - slot; if (q < 0) q += m; }
}

// One rotation of the register-resident columns x, y (cached squared norms nx, ny).
template <int R, int LG>
__device__ __forceinline__ bool rotate_regs(float (&x)[R], float (&y)[R], float& nx, float& ny, float tol2, unsigned gmask) {
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int e = 0; e < R; e += 2) {
        d0 = fmaf(x[e], y[e], d0);
        d1 = fmaf(x[e + 1], y[e + 1], d1);
    }
    const float xy = group_sum<LG>(d0 + d1, gmask);
    if (!(xy * xy > tol2 * nx * ny)) return false;
    float c, s, t;
    jacobi_cs(nx, ny, xy, c, s, t);
#pragma unroll
    for (int e = 0; e < R; ++e) {
        const float a = x[e], b = y[e];
        x[e] = c * a - s * b;
        y[e] = s * a + c * b;
    }
    nx -= t * xy;
    ny += t * xy;
    return true;
}

// KR     : rows per lane / 4 (column pitch LD = 32 KR + 4 floats)
// POLISH : second phase (needs a second n x LD buffer in shared memory)
// GLOBAL : working matrix in the global workspace instead of shared memory
template <int KR, bool POLISH, bool GLOBAL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
eigh_jacobi_kernel(const float* __restrict__ Ain, float* __restrict__ Wout, float* __restrict__ Uout,
                   int32_t* __restrict__ info, float* __restrict__ workspace,
                   int n, int canonical, float tol, int max_sweeps, int use_oddeven, int max_polish,
                   int redo_failed, long long ws_stride) {
    static_assert(!(POLISH && GLOBAL), "the polish phase needs both buffers in shared memory");
    // redo_failed: the fallback pass behind the tridiagonal-QL route -- only the graphs that route flagged with
    // info == -2 do any work; every other CTA leaves at once (uniform over the block: one word, read by all).
    if (redo_failed && reinterpret_cast<volatile int32_t*>(info)[blockIdx.x] != -2) return;
    constexpr int LD = 32 * KR + 4;
    constexpr int LG = kGroup;                                  // lanes per pair slot
    constexpr int VEC = 32 / LG;
    constexpr int R = VEC * KR;
    extern __shared__ __align__(16) float smem[];
    float* G;                                                   // [n][LD], column c at G + c*LD
    float* H;                                                   // [n][LD] (POLISH only)
    float* lam;                                                 // [n]
    if (GLOBAL) { G = workspace + (size_t)blockIdx.x * (size_t)ws_stride; H = nullptr; lam = smem; }
    else        { G = smem; H = G + (size_t)n * LD; lam = POLISH ? H + (size_t)n * LD : H; }
    float* sgn = lam + n;                                       // [n]
    float* nrm = sgn + n;                                       // [n]  cached squared column norms
    int* inv = reinterpret_cast<int*>(nrm + n);                 // [n]
    __shared__ float s_red[32];
    __shared__ float s_sigma;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = tid / LG, j = tid % LG;                       // pair slot, lane inside the slot
    const int ngroups = nthr / LG;
    const unsigned gmask = (LG == 8) ? (0xFFu << (lane & 24)) : (0xFFFFu << (lane & 16));   // the lanes of this slot
    const int n_pad = (n + 1) & ~1, npairs = n_pad >> 1, m = n_pad - 1;
    // odd-even transposition ordering (default): needs one pair slot per lane group and shared memory
    const bool oddeven = !GLOBAL && ngroups >= npairs && use_oddeven != 0;
    const float* A = Ain + (size_t)blockIdx.x * n * n;
    float* Wg = Wout + (size_t)blockIdx.x * n;
    float* Ug = Uout + (size_t)blockIdx.x * n * n;

    // ---- load: zero the padding, mirror the upper triangle, detect non-finite input.  The
    //      symmetrised matrix is also parked in the U output buffer (scratch until the final
    //      store) so that the polish can read its rows contiguously.
    for (int e = tid; e < n * LD; e += nthr) store_1<GLOBAL>(G + e, 0.f);
    __syncthreads();
    int bad = 0;
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, c = e - i * n;
        if (c >= i) {
            const float v = __ldg(A + e);
            if (!isfinite(v)) bad = 1;
            store_1<GLOBAL>(G + (size_t)i * LD + c, v);         // column i, row c  (= A_sym[c][i])
            if (c > i) store_1<GLOBAL>(G + (size_t)c * LD + i, v);
            if (POLISH) { Ug[e] = v; if (c > i) Ug[(size_t)c * n + i] = v; }
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
        for (int i = lane; i < n; i += 32) acc += fabsf(load_1<GLOBAL>(G + (size_t)c * LD + i));
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
    for (int c = tid; c < n; c += nthr) store_1<GLOBAL>(G + (size_t)c * LD + c, load_1<GLOBAL>(G + (size_t)c * LD + c) + sigma);
    __syncthreads();

    // A sweep that rotated at most this many pairs is the last one: what is left is second
    // order, and (POLISH) the next phase removes it anyway.
    const int few_threads = LG * (POLISH ? (n >> 3) : 0);

    // ---- phase 1: cyclic one-sided Jacobi sweeps on G
    int sweeps = 0, converged = 0, polish_sweeps = 0;
    while (sweeps < max_sweeps) {
        // exact squared norms once per sweep; inside the sweep they follow aa -= t ab, bb += t ab
        for (int k = g; k < n; k += ngroups) {
            float u[R];
            load_col<KR, GLOBAL, LG>(G + (size_t)k * LD, j, u);
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
            ss = group_sum<LG>(ss, gmask);
            if (j == 0) nrm[k] = ss;
        }
        __syncthreads();
        int rotated = 0;
        const float tol2 = tol * tol;
        if (oddeven) {
            // Odd-even transposition ordering: columns sit at POSITIONS 0..n-1; even steps pair
            // positions (2g, 2g+1), odd steps (2g+1, 2g+2); after every rotation the two columns
            // swap positions, so in n_pad steps every pair of columns meets exactly once.  Slot g
            // keeps the column at position 2g+1 in REGISTERS for the whole sweep and exchanges
            // only the other one through shared memory: it loads position 2g (even step) or 2g+2
            // (odd step), rotates, stores its previously kept column to that same position and
            // keeps the freshly loaded one.  Half the shared-memory traffic and half the LDS / STS
            // instructions of the round-robin formulation at the same number of resident warps.
            const int pk = 2 * g + 1;                                    // position of the kept column
            const bool have = g < npairs && pk < n;
            float ka[R], kb[R];                                            // the kept column alternates between the two
            float nka = 1.f, nkb = 1.f;
            if (have) { load_col<KR, GLOBAL, LG>(G + (size_t)pk * LD, j, ka); nka = nrm[pk]; }
            __syncthreads();
            for (int stp = 0; stp < n_pad; stp += 2) {
                // even step: exchange position 2g, kept column in ka -> afterwards kept in kb
                if (have) {
                    float* cx = G + (size_t)(2 * g) * LD;
                    load_col<KR, GLOBAL, LG>(cx, j, kb);
                    nkb = nrm[2 * g];
                    if (rotate_regs<R, LG>(kb, ka, nkb, nka, tol2, gmask)) rotated = 1;
                    store_col<KR, GLOBAL, LG>(cx, j, ka);
                    if (j == 0) nrm[2 * g] = nka;
                }
                __syncthreads();
                // odd step: exchange position 2g+2, kept column in kb -> afterwards kept in ka
                const bool odd_ok = have && (2 * g + 2 < n);
                if (odd_ok) {
                    float* cx = G + (size_t)(2 * g + 2) * LD;
                    load_col<KR, GLOBAL, LG>(cx, j, ka);
                    nka = nrm[2 * g + 2];
                    if (rotate_regs<R, LG>(kb, ka, nkb, nka, tol2, gmask)) rotated = 1;
                    store_col<KR, GLOBAL, LG>(cx, j, kb);
                    if (j == 0) nrm[2 * g + 2] = nkb;
                } else if (have) {
#pragma unroll
                    for (int e = 0; e < R; ++e) ka[e] = kb[e];                // idle end slot: the kept column just stays
                    nka = nkb;
                }
                __syncthreads();
            }
            if (have) { store_col<KR, GLOBAL, LG>(G + (size_t)pk * LD, j, ka); if (j == 0) nrm[pk] = nka; }
            __syncthreads();
        } else {
        for (int r = 0; r < m; ++r) {
            for (int slot = g; slot < npairs; slot += ngroups) {
                int p, q;
                rr_pair(slot, r, m, n_pad, p, q);
                if (p < n && q < n) {
                    float a[R], b[R];
                    load_col<KR, GLOBAL, LG>(G + (size_t)p * LD, j, a);
                    load_col<KR, GLOBAL, LG>(G + (size_t)q * LD, j, b);
                    float na = nrm[p], nb = nrm[q];
                    if (rotate_regs<R, LG>(a, b, na, nb, tol2, gmask)) {
                        store_col<KR, GLOBAL, LG>(G + (size_t)p * LD, j, a);
                        store_col<KR, GLOBAL, LG>(G + (size_t)q * LD, j, b);
                        if (j == 0) { nrm[p] = na; nrm[q] = nb; }
                        rotated = 1;
                    }
                }
            }
            __syncthreads();
        }
        }
        ++sweeps;
        if (__syncthreads_count(rotated) <= few_threads) { converged = 1; break; }
    }

    // ---- normalise the columns: U
    for (int k = g; k < n; k += ngroups) {
        float u[R];
        float* ck = G + (size_t)k * LD;
        load_col<KR, GLOBAL, LG>(ck, j, u);
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
        ss = group_sum<LG>(ss, gmask);
        const float inv_norm = 1.0f / sqrtf(ss);
#pragma unroll
        for (int e = 0; e < R; ++e) u[e] *= inv_norm;
        store_col<KR, GLOBAL, LG>(ck, j, u);
    }
    __syncthreads();

    if (POLISH) {
        // ---- phase 2a: H = A_sym U straight from the input (fresh, no accumulated rounding);
        //      each slot produces whole columns, lane j the rows it owns.
        const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ug) & 15u) == 0);
        for (int k = g; k < n; k += ngroups) {
            float h[R];
#pragma unroll
            for (int e = 0; e < R; ++e) h[e] = 0.f;
            const float* ck = G + (size_t)k * LD;
            for (int mm = 0; mm < n; ++mm) {
                const float u = ck[mm];
                const float* Ar = Ug + (size_t)mm * n;
#pragma unroll
                for (int kk = 0; kk < KR; ++kk) {
                    const int i0 = VEC * j + 32 * kk;
                    float av[VEC];
                    if (vec_ok && i0 + VEC - 1 < n) {
                        if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(Ar + i0); av[0] = t.x; av[1] = t.y; av[VEC - 2] = t.z; av[VEC - 1] = t.w; }
                        else { const float2 t = *reinterpret_cast<const float2*>(Ar + i0); av[0] = t.x; av[1] = t.y; }
                    } else {
#pragma unroll
                        for (int t = 0; t < VEC; ++t) av[t] = (i0 + t < n) ? Ar[i0 + t] : 0.f;
                    }
#pragma unroll
                    for (int t = 0; t < VEC; ++t) h[VEC * kk + t] = fmaf(av[t], u, h[VEC * kk + t]);
                }
            }
            store_col<KR, false, LG>(H + (size_t)k * LD, j, h);
        }
        __syncthreads();
        // ---- phase 2b: implicit two-sided Jacobi on T = U^T H (pivots are dot products of a
        //      U column with an H column; the rotation is applied to the columns of both).
        const float theta_tol = 2e-6f, noise = 3e-8f;
        for (int psweep = 0; psweep < max_polish; ++psweep) {
            ++polish_sweeps;
            int rotated = 0;
            for (int r = 0; r < m; ++r) {
                for (int slot = g; slot < npairs; slot += ngroups) {
                    int p, q;
                    rr_pair(slot, r, m, n_pad, p, q);
                    if (p < n && q < n) {
                        float up[R], uq[R], hx[R];
                        load_col<KR, false, LG>(G + (size_t)p * LD, j, up);
                        load_col<KR, false, LG>(G + (size_t)q * LD, j, uq);
                        load_col<KR, false, LG>(H + (size_t)p * LD, j, hx);            // h_p
                        float tpp = 0.f, tqp = 0.f, tqq = 0.f, tpq = 0.f;
#pragma unroll
                        for (int e = 0; e < R; ++e) { tpp = fmaf(up[e], hx[e], tpp); tqp = fmaf(uq[e], hx[e], tqp); }
                        load_col<KR, false, LG>(H + (size_t)q * LD, j, hx);            // h_q
#pragma unroll
                        for (int e = 0; e < R; ++e) { tqq = fmaf(uq[e], hx[e], tqq); tpq = fmaf(up[e], hx[e], tpq); }
                        tpp = group_sum<LG>(tpp, gmask); tqq = group_sum<LG>(tqq, gmask);
                        const float off = 0.5f * group_sum<LG>(tpq + tqp, gmask);
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
                            store_col<KR, false, LG>(G + (size_t)p * LD, j, up);
                            store_col<KR, false, LG>(G + (size_t)q * LD, j, uq);
                            load_col<KR, false, LG>(H + (size_t)p * LD, j, up);        // reuse registers: h_p
#pragma unroll
                            for (int e = 0; e < R; ++e) {
                                const float x = up[e], y = hx[e];
                                up[e] = c * x - s * y;
                                hx[e] = s * x + c * y;
                            }
                            store_col<KR, false, LG>(H + (size_t)p * LD, j, up);
                            store_col<KR, false, LG>(H + (size_t)q * LD, j, hx);
                            rotated = 1;
                        }
                    }
                }
                __syncthreads();
            }
            if (__syncthreads_count(rotated) <= few_threads) break;
        }
    }

    // ---- eigenvalues (Rayleigh quotients), unit columns, sign pivots
    for (int k = g; k < n; k += ngroups) {
        float u[R];
        float* ck = G + (size_t)k * LD;
        load_col<KR, GLOBAL, LG>(ck, j, u);
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < R; ++e) ss = fmaf(u[e], u[e], ss);
        ss = group_sum<LG>(ss, gmask);
        const float inv_norm = 1.0f / sqrtf(ss);
        float best = -1.f; int best_i = 0; float best_v = 0.f;
#pragma unroll
        for (int e = 0; e < R; ++e) {
            u[e] *= inv_norm;
            const int row = VEC * j + 32 * (e / VEC) + (e % VEC);
            const float mag = fabsf(u[e]);
            if (row < n && mag > best) { best = mag; best_i = row; best_v = u[e]; }
        }
        store_col<KR, GLOBAL, LG>(ck, j, u);
#pragma unroll
        for (int o = LG / 2; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(gmask, best, o);
            const int oi = __shfl_xor_sync(gmask, best_i, o);
            const float ov = __shfl_xor_sync(gmask, best_v, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_v = ov; }
        }
        if (j == 0) sgn[k] = (canonical && best_v < 0.f) ? -1.0f : 1.0f;
        float acc = 0.f;
        if (POLISH) {
            float hk[R];
            load_col<KR, false, LG>(H + (size_t)k * LD, j, hk);
#pragma unroll
            for (int e = 0; e < R; ++e) acc = fmaf(u[e], hk[e], acc);       // u normalised, h = A u_old
            acc *= inv_norm;                                                 // -> u^T A u
        } else {
            __syncwarp(gmask);
            // lambda_k = sum_i u_i (A_ii u_i + 2 sum_{mm > i} A[i][mm] u_mm) from the upper triangle
            for (int i = 0; i < n; ++i) {
                const float ui = load_1<GLOBAL>(ck + i);
                const float* Ar = A + (size_t)i * n;
                float part = 0.f;
#pragma unroll
                for (int kk = 0; kk < KR; ++kk) {
                    const int m0 = VEC * j + 32 * kk;
                    if (m0 + VEC - 1 < i) continue;
#pragma unroll
                    for (int t = 0; t < VEC; ++t) {
                        const float av = (m0 + t < n) ? __ldg(Ar + m0 + t) : 0.f;
                        const float f = (m0 + t > i) ? 2.f : ((m0 + t == i) ? 1.f : 0.f);
                        part = fmaf(av * f, u[VEC * kk + t], part);
                    }
                }
                acc = fmaf(ui, part, acc);
            }
        }
        acc = group_sum<LG>(acc, gmask);
        if (j == 0) lam[k] = acc;
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
        Ug[e] = load_1<GLOBAL>(G + (size_t)k * LD + i) * sgn[k];
    }
    if (tid == 0) info[blockIdx.x] = converged ? (sweeps | (polish_sweeps << 8)) : -2;
}

int g_eigh_max_polish = 1;      // gsr_set_tuning("eigh_max_polish", k): polish sweeps.  One is enough: it starts ~2e-3 from the
                                // solution and Jacobi converges quadratically there (measured: a 2nd sweep rotates 6-19 of 12720
                                // pairs and changes max |dU| vs float64 from 4.074e-4 to 4.084e-4, median 3.9e-5 both; costs 1.3 ms)
int g_eigh_oddeven = 1;         // gsr_set_tuning("eigh_oddeven", 0|1): 0 = round-robin ordering (19.6 ms / 1024 graphs), 1 = odd-even (14.6 ms)

static int eigh_threads(int n) {
    const int npairs = ((n + 1) & ~1) / 2;
    const int iters = (npairs * kGroup + 1023) / 1024;            // pair slots per group per round
    const int groups = (npairs + iters - 1) / iters;
    return 32 * ((groups * kGroup + 31) / 32);
}

template <int KR, bool POLISH, bool GLOBAL, int MAXT>
static int launch_one(const float* A, float* w, float* U, int32_t* info, float* ws, int batch, int n,
                      int canonical, float tol, int max_sweeps, size_t smem, cudaStream_t st, int redo, long long ws_stride) {
    auto k = eigh_jacobi_kernel<KR, POLISH, GLOBAL, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "eigh_jacobi: cudaFuncSetAttribute failed");
    const int threads = eigh_threads(n);
    if (threads > MAXT) return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: internal thread-count mismatch");
    k<<<batch, threads, smem, st>>>(A, w, U, info, ws, n, canonical, tol, max_sweeps, g_eigh_oddeven, g_eigh_max_polish,
                                    redo, ws_stride);
    return launched(redo ? "eigh_jacobi_redo" : "eigh_jacobi");
}

template <int KR>
static int launch_eigh(const float* A, float* w, float* U, int32_t* info, float* ws, int batch, int n,
                       int canonical, float tol, int max_sweeps, cudaStream_t st, int redo = 0, long long ws_stride = 0) {
    constexpr int LD = 32 * KR + 4;
    if (ws_stride == 0) ws_stride = (long long)n * LD;
    constexpr int MAXT = (128 * KR < 1024) ? 128 * KR : (KR >= 9 ? 576 : 1024);
    const size_t small = 4 * (size_t)n * sizeof(float);
    const size_t one = (size_t)n * LD * sizeof(float) + small, two = 2 * (size_t)n * LD * sizeof(float) + small;
    const size_t cap = 227 * 1024;
    if (two <= cap) return launch_one<KR, true, false, MAXT>(A, w, U, info, ws, batch, n, canonical, tol, max_sweeps, two, st, redo, ws_stride);
    if (one <= cap) return launch_one<KR, false, false, MAXT>(A, w, U, info, ws, batch, n, canonical, tol, max_sweeps, one, st, redo, ws_stride);
    if (!ws) return fail(GSR_E_ARG, "eigh_jacobi: this n needs a global workspace (gsr_eigh_workspace)");
    if (ws_stride < (long long)n * LD) return fail(GSR_E_ARG, "eigh_jacobi: workspace stride below n * (32 ceil(n / 32) + 4) floats");
    return launch_one<KR, false, true, MAXT>(A, w, U, info, ws, batch, n, canonical, tol, max_sweeps, small, st, redo, ws_stride);
}

static int eigh_dispatch(const float* A, float* w, float* U, int32_t* info, float* workspace, int batch, int n,
                         int canonical_sign, float tol, int max_sweeps, cudaStream_t st, int redo, long long ws_stride) {
#define GSR_EIGH_CASE(KR) case KR: return launch_eigh<KR>(A, w, U, info, workspace, batch, n, canonical_sign, tol, max_sweeps, st, redo, ws_stride)
    switch ((n + 31) / 32) {
        GSR_EIGH_CASE(1); GSR_EIGH_CASE(2); GSR_EIGH_CASE(3); GSR_EIGH_CASE(4); GSR_EIGH_CASE(5);
        GSR_EIGH_CASE(6); GSR_EIGH_CASE(7); GSR_EIGH_CASE(8); GSR_EIGH_CASE(9);
        default: return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: unsupported n");
    }
#undef GSR_EIGH_CASE
}

// Device-side fallback of the tridiagonal-QL route (csrc/eigh_tridiag.cu): the graphs it flagged with info == -2 are
// redone by the Jacobi kernel in the same stream, no host round trip.  `scratch` (floats, `stride` per graph) is only
// touched for n whose matrix does not fit shared memory.
int eigh_jacobi_redo_failed(const float* A, float* w, float* U, int32_t* info, float* scratch, long long stride,
                            int batch, int n, int canonical_sign, cudaStream_t st) {
    return eigh_dispatch(A, w, U, info, scratch, batch, n, canonical_sign, 1e-6f, 30, st, 1, stride);
}

}  // namespace gsr

using namespace gsr;

extern "C" size_t gsr_eigh_workspace(int batch, int n) {
    if (n <= 0 || batch <= 0 || n > GSR_EIGH_MAX_N) return 0;
    const size_t LD = 32 * (size_t)((n + 31) / 32) + 4;
    const size_t one = (size_t)n * LD * sizeof(float) + 4 * (size_t)n * sizeof(float);
    return one <= 227 * 1024 ? 0 : (size_t)batch * n * LD;
}

extern "C" int gsr_eigh_jacobi_f32(const float* A, float* w, float* U, int32_t* info, float* workspace,
                                   int batch, int n, int canonical_sign, float tol, int max_sweeps,
                                   void* stream) {
    if (!A || !w || !U || !info || batch < 0 || n <= 0) return fail(GSR_E_ARG, "eigh_jacobi: bad argument");
    if (batch == 0) return 0;
    if (tol <= 0.f) tol = 1e-6f;
    if (max_sweeps <= 0) max_sweeps = 30;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > GSR_EIGH_MAX_N) return fail(GSR_E_UNSUPPORTED, "eigh_jacobi: n > 288 is not supported");
    return eigh_dispatch(A, w, U, info, workspace, batch, n, canonical_sign, tol, max_sweeps, st, 0, 0);
}
