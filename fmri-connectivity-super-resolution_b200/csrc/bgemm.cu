// Batched small-GEMM family of the GSR-Net path, SIMT FP32 (FFMA) variant:
//   models/layers.py:37,45,54,59,85-86, models/ops.py:60 and their autograd.
//
// Tile shape is chosen for the problem, not for powers of two: the HR / hidden
// dimension is 268 = 2 * 134, so a 136 x 136 CTA tile (17 x 17 threads, 8 x 8
// register micro-tile split into two 4-wide halves) covers a 268 x 268 subject
// matrix with four CTAs at 97 % tile efficiency (a 128-wide tile would waste
// 30 %).  Operand tiles are staged k-major in shared memory (pitch 140 floats)
// whatever the global layout, so the inner loop is one code path: per k step
// four 128-bit shared loads feed 64 FFMAs.  Global loads are register-staged
// (next k-tile is in flight while the current one is multiplied), one
// __syncthreads per k-tile, two shared buffers.
//
// Shared operands (weights) use a zero batch stride.  Weight gradients use the
// "reduce" entry: subjects are dealt round-robin to `splits` CTAs per tile, each
// accumulates its subjects in registers, partial tiles are summed in a fixed
// order by a second kernel (deterministic, no atomics).
#include "common.cuh"
#include <string.h>

namespace gsr {

constexpr int kT = 17;               // threads per tile edge
constexpr int kBM = 8 * kT;          // 136
constexpr int kBN = 8 * kT;          // 136
constexpr int kHalf = 4 * kT;        // 68
constexpr int kPitch = 140;          // floats per k row in shared memory
constexpr int kThreads = 320;        // 289 compute threads, all 320 load

// Tuning variants, selected at run time with gsr_set_tuning("gemm_variant", v):
//   0: BK = 16, one CTA per SM  (151 registers, no spills, 10 warps / SM)  <- default: fastest
//      on B200 (32.6 TFLOP/s alone, 28.8 in the step; variant 1: 31.9 / 26.1, variant 2: 28.5)
//   1: BK =  8, two CTAs per SM (96 registers, a few spills, 20 warps / SM)
//   2: BK =  8, one CTA per SM
int g_gemm_variant = 0;
int g_lin_small = 1;      // one-subject products on lin::lin_kernel (gsr_set_tuning("bgemm_lin", 0): the 136 x 136 kernel for everything)
extern int g_eigh_oddeven;
extern int g_eigh_max_polish;
extern int g_eigh_ql_max_iter;
extern int g_eigh_redo;
extern int g_eigh_k2_lanes;
extern int g_eigh_chunks;
extern int g_eigh_k2_warps;
extern int g_eigh_k1_regblocks;
extern int g_norm_bwd_cluster;
extern int g_eigh_q_launch;
namespace tc { extern int g_pairs, g_tail, g_dual, g_tma, g_utccp, g_mn_partial, g_pf_ahead, g_tmem_a, g_cluster_release, g_partial_head, g_reduce_simt_tails; }

struct GemmParams {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    long long sA, sB, sC;
    int batch, splits;
    float alpha;
    int epilogue;
    const float* extra; int ld_extra; long long s_extra;
};

// Load one float4 chunk of an operand tile.
//   KCONTIG: the operand is stored [row][k] (k contiguous)  -> chunk = (row r, 4 consecutive k)
//   else   : the operand is stored [k][row] (row contiguous) -> chunk = (k, 4 consecutive rows)
template <int BK, bool KCONTIG, bool VEC>
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ base, int ld, int rows, int K,
                                             int row0, int k0, int c) {
    constexpr int kChunks = kBM * BK / 4;
    constexpr int kQ = BK / 4;           // float4 chunks per row when k is contiguous
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= kChunks) return v;
    if (KCONTIG) {
        const int r = row0 + c / kQ, k = k0 + ((c % kQ) << 2);
        if (r < rows && k < K) {
            const float* p = base + (size_t)r * ld + k;
            if (VEC && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
            else {
                v.x = __ldg(p);
                if (k + 1 < K) v.y = __ldg(p + 1);
                if (k + 2 < K) v.z = __ldg(p + 2);
                if (k + 3 < K) v.w = __ldg(p + 3);
            }
        }
    } else {
        const int k = k0 + c / (kBM / 4), r = row0 + ((c % (kBM / 4)) << 2);
        if (k < K && r < rows) {
            const float* p = base + (size_t)k * ld + r;
            if (VEC && r + 3 < rows) v = __ldg(reinterpret_cast<const float4*>(p));
            else {
                v.x = __ldg(p);
                if (r + 1 < rows) v.y = __ldg(p + 1);
                if (r + 2 < rows) v.z = __ldg(p + 2);
                if (r + 3 < rows) v.w = __ldg(p + 3);
            }
        }
    }
    return v;
}

template <int BK, bool KCONTIG>
__device__ __forceinline__ void store_chunk(float* __restrict__ s, int c, const float4& v) {
    constexpr int kChunks = kBM * BK / 4;
    constexpr int kQ = BK / 4;
    if (c >= kChunks) return;
    if (KCONTIG) {
        const int r = c / kQ, k = (c % kQ) << 2;
        s[(k + 0) * kPitch + r] = v.x;
        s[(k + 1) * kPitch + r] = v.y;
        s[(k + 2) * kPitch + r] = v.z;
        s[(k + 3) * kPitch + r] = v.w;
    } else {
        const int k = c / (kBM / 4), r = (c % (kBM / 4)) << 2;
        *reinterpret_cast<float4*>(s + k * kPitch + r) = v;
    }
}

// TA: op(A) = A^T (A stored [K][M]);  TB: op(B) = B^T (B stored [N][K]).
template <int BK, int OCC, bool TA, bool TB, bool VEC, bool REDUCE>
__global__ void __launch_bounds__(kThreads, OCC)
bgemm_kernel(const GemmParams p) {
    constexpr int kBK = BK;
    constexpr int kChunksPerThread = (kBM * BK / 4 + kThreads - 1) / kThreads;
    __shared__ __align__(16) float As[2][kBK * kPitch];
    __shared__ __align__(16) float Bs[2][kBK * kPitch];
    const int tid = threadIdx.x;
    const int ty = tid / kT, tx = tid - ty * kT;
    const bool compute = tid < kT * kT;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + kBK - 1) / kBK;
    const int b_first = blockIdx.z;
    const int b_step = REDUCE ? p.splits : p.batch;     // non-reduce: exactly one subject per CTA

    for (int b = b_first; b < p.batch; b += b_step) {
        const float* A = p.A + (size_t)b * p.sA;
        const float* B = p.B + (size_t)b * p.sB;
        float4 ra[kChunksPerThread], rb[kChunksPerThread];
#pragma unroll
        for (int i = 0; i < kChunksPerThread; ++i) {
            ra[i] = load_chunk<BK, !TA, VEC>(A, p.lda, p.M, p.K, m0, 0, tid + i * kThreads);
            rb[i] = load_chunk<BK, TB, VEC>(B, p.ldb, p.N, p.K, n0, 0, tid + i * kThreads);
        }
        __syncthreads();           // previous subject's last tile fully consumed (REDUCE loop)
#pragma unroll
        for (int i = 0; i < kChunksPerThread; ++i) {
            store_chunk<BK, !TA>(As[0], tid + i * kThreads, ra[i]);
            store_chunk<BK, TB>(Bs[0], tid + i * kThreads, rb[i]);
        }
        __syncthreads();
        for (int kt = 0; kt < nk; ++kt) {
            const int cur = kt & 1;
            if (kt + 1 < nk) {
#pragma unroll
                for (int i = 0; i < kChunksPerThread; ++i) {
                    ra[i] = load_chunk<BK, !TA, VEC>(A, p.lda, p.M, p.K, m0, (kt + 1) * kBK, tid + i * kThreads);
                    rb[i] = load_chunk<BK, TB, VEC>(B, p.ldb, p.N, p.K, n0, (kt + 1) * kBK, tid + i * kThreads);
                }
            }
            if (compute) {
                const float* as = As[cur] + ty * 4;
                const float* bs = Bs[cur] + tx * 4;
#pragma unroll
                for (int k = 0; k < kBK; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4*>(as + k * kPitch);
                    const float4 a1 = *reinterpret_cast<const float4*>(as + k * kPitch + kHalf);
                    const float4 b0 = *reinterpret_cast<const float4*>(bs + k * kPitch);
                    const float4 b1 = *reinterpret_cast<const float4*>(bs + k * kPitch + kHalf);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                }
            }
            if (kt + 1 < nk) {
#pragma unroll
                for (int i = 0; i < kChunksPerThread; ++i) {
                    store_chunk<BK, !TA>(As[cur ^ 1], tid + i * kThreads, ra[i]);
                    store_chunk<BK, TB>(Bs[cur ^ 1], tid + i * kThreads, rb[i]);
                }
            }
            __syncthreads();
        }
    }
    if (!compute) return;

    // ---- epilogue
    float* C;
    int ldc;
    if (REDUCE) { C = p.C + (size_t)blockIdx.z * p.M * p.N; ldc = p.N; }
    else        { C = p.C + (size_t)blockIdx.z * p.sC;      ldc = p.ldc; }
    const float* extra = nullptr;
    if (!REDUCE && p.epilogue == GSR_EPI_ADD) extra = p.extra + (size_t)blockIdx.z * p.s_extra;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : kHalf + ty * 4 + (i - 4));
        if (m >= p.M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * kHalf + tx * 4;
            if (n >= p.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][h * 4 + j];
                if (!REDUCE) {
                    x *= p.alpha;
                    if (n + j < p.N) {
                        if (p.epilogue == GSR_EPI_BIAS) x += __ldg(p.extra + n + j);
                        else if (p.epilogue == GSR_EPI_ADD) x += __ldg(extra + (size_t)m * p.ld_extra + n + j);
                    }
                }
                v[j] = x;
            }
            float* dst = C + (size_t)m * ldc + n;
            if (VEC && n + 3 < p.N && ((ldc & 3) == 0)) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < p.N) dst[j] = v[j];
            }
        }
    }
}

// C[m][n] = alpha * sum_z ws[z][m][n] (+ extra[m][n]); fixed summation order.
__global__ void reduce_partials_kernel(const float* __restrict__ ws, float* __restrict__ C,
                                       int M, int N, int ldc, int splits, float alpha,
                                       const float* __restrict__ extra, int ld_extra) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * N) return;
    const int m = e / N, n = e - m * N;
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += __ldg(ws + (size_t)z * M * N + e);
    acc *= alpha;
    if (extra) acc += __ldg(extra + (size_t)m * ld_extra + n);
    C[(size_t)m * ldc + n] = acc;
}

// Column sums of X[rows][cols] (Linear bias gradient): 32 columns x 8 row phases per CTA, the eight partial sums of a
// column folded in a fixed order (a thread per column walking all rows was a chain of `rows` dependent loads: 9-16 us
// for the 160 x 320 gradients of the Graph U-Net, which sit on the tail of the step).
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int rows, int cols, int ld) {
    __shared__ float part[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + x;
    float acc = 0.f;
    if (c < cols)
        for (int r = y; r < rows; r += 8) acc += __ldg(X + (size_t)r * ld + c);
    part[y][x] = acc;
    __syncthreads();
    if (y == 0 && c < cols) {
        float s = part[0][x];
#pragma unroll
        for (int i = 1; i < 8; ++i) s += part[i][x];
        out[c] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// One-subject products of the Linear-layer class (models/ops.py:60 and its autograd: the Graph U-Net runs once per step
// on <= 160 x 640 operands; also the reference's batch-1 calls).  These are latency problems, not throughput problems:
// the 136 x 136 kernel above gives a 160 x 320 x 320 product 6 CTAs walking 20 k-tiles with one global-load round trip
// each (42-48 us per launch, 21 launches per step, the backward ones on the step's tail).  Here a CTA owns a 32 x 32
// tile of C, brings its WHOLE operand panels (32 x K of op(A), K x 32 of op(B)) into shared memory with one burst of
// 16-byte cp.async (zero filled at the ragged edges by the copy's source size) in the orientation they have in global
// memory, and its four 64-thread groups each take every fourth 4-k chunk with a 4 x 4 register tile; the four partial
// tiles meet in shared memory in a fixed order (deterministic, no atomics).  160 x 320: 50 CTAs, one load round trip.
namespace lin {

constexpr int TM = 32, TN = 32, THREADS = 256, KG = 4;
constexpr int PR = 36;                        // pitch of a [k][32 rows] panel image (16-byte aligned rows)
constexpr int MAXK = 768;                     // two panels of 768 x 36 floats = 216 KB
constexpr int RED_FLOATS = KG * TM * 33;

__host__ __device__ inline int pitch_k(int K) { return ((K + 31) & ~31) + 4; }      // = 4 mod 32: rows of a [32][k] panel start one 16-byte bank group apart
__host__ __device__ inline int panel_floats(bool kcontig, int K) { return kcontig ? TM * pitch_k(K) : ((K + 3) & ~3) * PR; }

__device__ __forceinline__ void cp16z(float* dst, const float* src, int bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
}

// KC: the operand is stored [row][k] (k contiguous) -> image [32][pitch_k]; else stored [k][row] -> image [k][PR].
template <bool KC>
__device__ __forceinline__ void stage(float* s, const float* __restrict__ g, int ld, int rows, int K, int row0, int PK, int tid) {
    if (KC) {
        const int cpr = (K + 3) >> 2;
        for (int c = tid; c < TM * cpr; c += THREADS) {
            const int r = c / cpr, k = (c - r * cpr) << 2;
            const int bytes = (row0 + r < rows) ? min(16, 4 * (K - k)) : 0;
            cp16z(s + r * PK + k, bytes ? g + (size_t)(row0 + r) * ld + k : g, bytes);
        }
    } else {
        const int kr = (K + 3) & ~3;                                   // rows K .. kr-1 of the image are read by the last chunk: zero
        for (int c = tid; c < 8 * kr; c += THREADS) {
            const int k = c >> 3, rc = (c & 7) << 2;
            const int left = (k < K) ? rows - row0 - rc : 0;
            const int bytes = left > 0 ? min(16, 4 * left) : 0;
            cp16z(s + k * PR + rc, bytes ? g + (size_t)k * ld + row0 + rc : g, bytes);
        }
    }
}

// TA: op(A) = A^T (A stored [K][M]);  TB: op(B) = B^T (B stored [N][K]).
template <bool TA, bool TB>
__global__ void __launch_bounds__(THREADS) lin_kernel(const GemmParams p) {
    extern __shared__ __align__(16) float lsm[];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, PK = pitch_k(K);
    float* As = lsm;
    float* Bs = lsm + panel_floats(!TA, K);
    stage<!TA>(As, p.A, p.lda, p.M, K, m0, PK, tid);
    stage<TB>(Bs, p.B, p.ldb, p.N, K, n0, PK, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // thread (tm, tn) of group kg: rows 4 tm + i; columns 4 tn + j of a [k][n] image, tn + 8 j of a [n][k] image (there the
    // eight lanes of a quarter warp read eight different rows: one bank group apart only when the rows are consecutive)
    const int kg = tid >> 6, tm = (tid & 63) >> 3, tn = tid & 7;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int nkc = (K + 3) >> 2;
    for (int kc = kg; kc < nkc; kc += KG) {
        float a[4][4], b[4][4];                                        // [k within the chunk][row | column]
        if (!TA) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(As + (4 * tm + i) * PK + 4 * kc);
                a[0][i] = v.x; a[1][i] = v.y; a[2][i] = v.z; a[3][i] = v.w;
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v = *reinterpret_cast<const float4*>(As + (4 * kc + kk) * PR + 4 * tm);
                a[kk][0] = v.x; a[kk][1] = v.y; a[kk][2] = v.z; a[kk][3] = v.w;
            }
        }
        if (TB) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + (tn + 8 * j) * PK + 4 * kc);
                b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + (4 * kc + kk) * PR + 4 * tn);
                b[kk][0] = v.x; b[kk][1] = v.y; b[kk][2] = v.z; b[kk][3] = v.w;
            }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[kk][i], b[kk][j], acc[i][j]);
    }

    __syncthreads();                                                   // every group is done with the panels
    float* red = lsm;                                                  // [KG][32][33]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(kg * TM + 4 * tm + i) * 33 + (TB ? tn + 8 * j : 4 * tn + j)] = acc[i][j];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TM * TN / THREADS; ++q) {
        const int e = tid + THREADS * q, ml = e >> 5, nl = e & 31;
        const int m = m0 + ml, n = n0 + nl;
        if (m >= p.M || n >= p.N) continue;
        float x = red[ml * 33 + nl];
#pragma unroll
        for (int g = 1; g < KG; ++g) x += red[(g * TM + ml) * 33 + nl];
        x *= p.alpha;
        if (p.epilogue == GSR_EPI_BIAS) x += __ldg(p.extra + n);
        else if (p.epilogue == GSR_EPI_ADD) x += __ldg(p.extra + (size_t)m * p.ld_extra + n);
        p.C[(size_t)m * p.ldc + n] = x;
    }
}

static size_t smem_bytes(int transA, int transB, int K) {
    const size_t f = (size_t)panel_floats(!transA, K) + (size_t)panel_floats(transB != 0, K);
    return sizeof(float) * (f > (size_t)RED_FLOATS ? f : (size_t)RED_FLOATS);
}

static bool eligible(const GemmParams& p) {
    return p.batch == 1 && p.K >= 1 && p.K <= MAXK && (long long)p.M * p.N <= (1LL << 18) && aligned16(p.A) && aligned16(p.B) &&
           (p.lda % 4 == 0) && (p.ldb % 4 == 0);
}

static int launch(const GemmParams& p, int transA, int transB, cudaStream_t st) {
    static const bool attr_ok = [] {
        const int mx = (int)(sizeof(float) * 2 * (size_t)(((MAXK + 3) & ~3) * PR));
        return cudaFuncSetAttribute(lin_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) == cudaSuccess &&
               cudaFuncSetAttribute(lin_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) == cudaSuccess &&
               cudaFuncSetAttribute(lin_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) == cudaSuccess &&
               cudaFuncSetAttribute(lin_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) == cudaSuccess;
    }();
    if (!attr_ok) { cudaGetLastError(); return fail(GSR_E_UNSUPPORTED, "bgemm: shared-memory opt-in of the one-subject kernel failed"); }
    const dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, 1);
    const size_t smem = smem_bytes(transA, transB, p.K);
    if (!transA && !transB) lin_kernel<false, false><<<grid, THREADS, smem, st>>>(p);
    else if (!transA && transB) lin_kernel<false, true><<<grid, THREADS, smem, st>>>(p);
    else if (transA && !transB) lin_kernel<true, false><<<grid, THREADS, smem, st>>>(p);
    else lin_kernel<true, true><<<grid, THREADS, smem, st>>>(p);
    return launched("bgemm_lin");
}

}  // namespace lin

template <int BK, int OCC, bool REDUCE>
static void launch_variant(const GemmParams& p, int transA, int transB, bool vec, dim3 grid, cudaStream_t st) {
#define GSR_LAUNCH(TA, TB, V) bgemm_kernel<BK, OCC, TA, TB, V, REDUCE><<<grid, kThreads, 0, st>>>(p)
    if (vec) {
        if (!transA && !transB) GSR_LAUNCH(false, false, true);
        else if (!transA && transB) GSR_LAUNCH(false, true, true);
        else if (transA && !transB) GSR_LAUNCH(true, false, true);
        else GSR_LAUNCH(true, true, true);
    } else {
        if (!transA && !transB) GSR_LAUNCH(false, false, false);
        else if (!transA && transB) GSR_LAUNCH(false, true, false);
        else if (transA && !transB) GSR_LAUNCH(true, false, false);
        else GSR_LAUNCH(true, true, false);
    }
#undef GSR_LAUNCH
}

template <bool REDUCE>
static int launch_gemm(const GemmParams& p, int transA, int transB, bool vec, dim3 grid, cudaStream_t st) {
    switch (g_gemm_variant) {
        case 1: launch_variant<8, 2, REDUCE>(p, transA, transB, vec, grid, st); break;
        case 2: launch_variant<8, 1, REDUCE>(p, transA, transB, vec, grid, st); break;
        default: launch_variant<16, 1, REDUCE>(p, transA, transB, vec, grid, st); break;
    }
    return launched(REDUCE ? "bgemm_reduce" : "bgemm");
}

static bool vec_ok(const GemmParams& p) {
    return aligned16(p.A) && aligned16(p.B) && aligned16(p.C) && (p.lda % 4 == 0) && (p.ldb % 4 == 0) &&
           (p.sA % 4 == 0) && (p.sB % 4 == 0);
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_bgemm_f32(const float* A, const float* B, float* C,
                             int M, int N, int K, int lda, int ldb, int ldc,
                             long long strideA, long long strideB, long long strideC,
                             int transA, int transB, int batch, float alpha,
                             int epilogue, const float* extra, int ld_extra, long long stride_extra,
                             void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K < 0 || batch < 0) return fail(GSR_E_ARG, "bgemm: bad argument");
    if (epilogue != GSR_EPI_NONE && !extra) return fail(GSR_E_ARG, "bgemm: epilogue needs `extra`");
    if (epilogue < 0 || epilogue > GSR_EPI_ADD) return fail(GSR_E_ARG, "bgemm: unknown epilogue");
    if (M == 0 || N == 0 || batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "bgemm: batch > 65535");
    GemmParams p{A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, 1, alpha,
                 epilogue, extra, ld_extra, stride_extra};
    if (g_lin_small && lin::eligible(p)) return lin::launch(p, transA, transB, (cudaStream_t)stream);
    dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM, batch);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "bgemm: M too large");
    const bool vec = vec_ok(p) && (ldc % 4 == 0) && (strideC % 4 == 0);
    return launch_gemm<false>(p, transA, transB, vec, grid, (cudaStream_t)stream);
}

extern "C" int gsr_set_tuning(const char* key, int value) {
    if (!key) return fail(GSR_E_ARG, "set_tuning: null key");
    if (!strcmp(key, "gemm_variant")) { gsr::g_gemm_variant = value; return 0; }
    if (!strcmp(key, "bgemm_lin")) { gsr::g_lin_small = value; return 0; }
    if (!strcmp(key, "eigh_oddeven")) { gsr::g_eigh_oddeven = value; return 0; }
    if (!strcmp(key, "eigh_max_polish")) { gsr::g_eigh_max_polish = value; return 0; }
    if (!strcmp(key, "eigh_ql_max_iter")) { gsr::g_eigh_ql_max_iter = value; return 0; }
    if (!strcmp(key, "eigh_k2_warps")) { gsr::g_eigh_k2_warps = value; return 0; }
    if (!strcmp(key, "eigh_k1_regblocks")) { gsr::g_eigh_k1_regblocks = value; return 0; }
    if (!strcmp(key, "eigh_q_launch")) { gsr::g_eigh_q_launch = value; return 0; }
    if (!strcmp(key, "normalize_bwd_cluster")) { gsr::g_norm_bwd_cluster = value; return 0; }
    if (!strcmp(key, "eigh_chunks")) { gsr::g_eigh_chunks = value; return 0; }
    if (!strcmp(key, "eigh_k2_lanes")) { gsr::g_eigh_k2_lanes = value; return 0; }
    if (!strcmp(key, "eigh_redo_failed")) { gsr::g_eigh_redo = value; return 0; }
    if (!strcmp(key, "tcgemm_pairs")) { gsr::tc::g_pairs = value; return 0; }
    if (!strcmp(key, "tcgemm_ragged_rows")) { gsr::tc::g_tail = value; return 0; }
    if (!strcmp(key, "tcgemm_tma")) { gsr::tc::g_tma = value; return 0; }
    if (!strcmp(key, "tcgemm_utccp")) { gsr::tc::g_utccp = value; return 0; }
    if (!strcmp(key, "tcgemm_tmem_a")) { gsr::tc::g_tmem_a = value; return 0; }
    if (!strcmp(key, "tcgemm_cluster_release")) { gsr::tc::g_cluster_release = value; return 0; }
    if (!strcmp(key, "tcgemm_partial_head")) { gsr::tc::g_partial_head = value; return 0; }
    if (!strcmp(key, "tcgemm_reduce_simt_tails")) { gsr::tc::g_reduce_simt_tails = value; return 0; }
    if (!strcmp(key, "tcgemm_prefetch")) { gsr::tc::g_pf_ahead = value < 0 ? 0 : value; return 0; }
    if (!strcmp(key, "tcgemm_dual")) { gsr::tc::g_dual = value; return 0; }
    if (!strcmp(key, "tcgemm_mn_partial")) { gsr::tc::g_mn_partial = value; return 0; }
    return fail(GSR_E_ARG, "set_tuning: unknown key");
}

extern "C" size_t gsr_bgemm_reduce_workspace(int M, int N, int batch, int* splits_out) {
    const int tiles = ((N + kBN - 1) / kBN) * ((M + kBM - 1) / kBM);
    int splits = (2 * kNumSMs + tiles - 1) / tiles;      // about one full wave at 2 CTAs / SM
    if (splits > batch) splits = batch;
    if (splits < 1) splits = 1;
    if (splits_out) *splits_out = splits;
    return (size_t)splits * (size_t)M * (size_t)N;
}

extern "C" int gsr_bgemm_reduce_f32(const float* A, const float* B, float* C,
                                    int M, int N, int K, int lda, int ldb, int ldc,
                                    long long strideA, long long strideB,
                                    int transA, int transB, int batch, float alpha,
                                    int epilogue, const float* extra, int ld_extra,
                                    float* workspace, int splits, void* stream) {
    if (!A || !B || !C || !workspace || M <= 0 || N <= 0 || K < 0 || batch <= 0 || splits <= 0)
        return fail(GSR_E_ARG, "bgemm_reduce: bad argument");
    if (epilogue != GSR_EPI_NONE && epilogue != GSR_EPI_ADD) return fail(GSR_E_ARG, "bgemm_reduce: epilogue must be NONE or ADD");
    if (epilogue == GSR_EPI_ADD && !extra) return fail(GSR_E_ARG, "bgemm_reduce: epilogue needs `extra`");
    if (splits > batch) splits = batch;
    GemmParams p{A, B, workspace, M, N, K, lda, ldb, N, strideA, strideB, 0, batch, splits, 1.0f,
                 GSR_EPI_NONE, nullptr, 0, 0};
    dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM, splits);
    const bool vec = vec_ok(p) && aligned16(workspace) && (N % 4 == 0) && (((long long)M * N) % 4 == 0);
    int rc = launch_gemm<true>(p, transA, transB, vec, grid, (cudaStream_t)stream);
    if (rc) return rc;
    const int total = M * N;
    reduce_partials_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        workspace, C, M, N, ldc, splits, alpha, epilogue == GSR_EPI_ADD ? extra : nullptr, ld_extra);
    return launched("bgemm_reduce_partials");
}

// Sum of the `splits` partial products of a K-split contraction (workspace [splits][M][N], fixed order), scaled, with the
// epilogue of the product: + bias[n] or + extra[m][n] (extra may be C itself: one thread reads and writes an element).
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* C, int M, int N, int ldc, int splits, float alpha,
                                     int epilogue, const float* extra, int ld_extra) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x, total = (long long)M * N;
    if (e >= total) return;
    const int m = (int)(e / N), n = (int)(e - (long long)m * N);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += __ldg(ws + (size_t)z * total + e);
    acc *= alpha;
    if (epilogue == GSR_EPI_BIAS) acc += extra[n];
    else if (epilogue == GSR_EPI_ADD) acc += extra[(size_t)m * ld_extra + n];
    C[(size_t)m * ldc + n] = acc;
}

extern "C" int gsr_splitk_reduce_f32(const float* workspace, float* C, int M, int N, int ldc, int splits, float alpha,
                                     int epilogue, const float* extra, int ld_extra, void* stream) {
    if (!workspace || !C || M <= 0 || N <= 0 || ldc < N || splits <= 0) return fail(GSR_E_ARG, "splitk_reduce: bad argument");
    if (epilogue < 0 || epilogue > GSR_EPI_ADD || (epilogue != GSR_EPI_NONE && !extra)) return fail(GSR_E_ARG, "splitk_reduce: bad epilogue");
    const long long total = (long long)M * N;
    if ((total + 255) / 256 > 0x7fffffffLL) return fail(GSR_E_UNSUPPORTED, "splitk_reduce: M * N too large");
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(workspace, C, M, N, ldc, splits, alpha,
                                                                                          epilogue, extra, ld_extra);
    return launched("splitk_reduce");
}

extern "C" int gsr_colsum_f32(const float* X, float* out, int rows, int cols, int ld, void* stream) {
    if (!X || !out || rows < 0 || cols <= 0) return fail(GSR_E_ARG, "colsum: bad argument");
    colsum_kernel<<<(cols + 31) / 32, 256, 0, (cudaStream_t)stream>>>(X, out, rows, cols, ld);
    return launched("colsum");
}
