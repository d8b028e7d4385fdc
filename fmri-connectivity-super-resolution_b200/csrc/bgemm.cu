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

// Column sums of X[rows][cols] (Linear bias gradient); one thread per column, fixed order.
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int rows, int cols, int ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += __ldg(X + (size_t)r * ld + c);
    out[c] = acc;
}

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
    dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM, batch);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "bgemm: M too large");
    const bool vec = vec_ok(p) && (ldc % 4 == 0) && (strideC % 4 == 0);
    return launch_gemm<false>(p, transA, transB, vec, grid, (cudaStream_t)stream);
}

extern "C" int gsr_set_tuning(const char* key, int value) {
    if (!key) return fail(GSR_E_ARG, "set_tuning: null key");
    if (!strcmp(key, "gemm_variant")) { gsr::g_gemm_variant = value; return 0; }
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

extern "C" int gsr_colsum_f32(const float* X, float* out, int rows, int cols, int ld, void* stream) {
    if (!X || !out || rows < 0 || cols <= 0) return fail(GSR_E_ARG, "colsum: bad argument");
    colsum_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(X, out, rows, cols, ld);
    return launched("colsum");
}
