// Batched symmetric eigendecomposition, tridiagonal route (n <= 160): the in-step LR
// decomposition of the GSR layer, torch.linalg.eigh(A, UPLO='U') of models/layers.py:27.
//
// The one-sided Jacobi kernel (eigh.cu) executes ~18 n^3 flops per graph (8 sweeps + polish) and
// every one of them moves its operands through shared memory: it is bound by the LSU /
// shared-memory pipe.  This path does the classical three-stage reduction in ~6 n^3 flops and
// keeps the MATRICES IN REGISTERS: one thread owns one row (160 fp32 registers of the 64 K per
// SM; 160 threads per graph, two graphs per SM), shared memory only streams the operands that
// are uniform over the warp (the Householder vector, the plane rotations), one broadcast
// wavefront per four columns.
//
//   K1  tridiag_kernel   Householder reduction, last column first (LAPACK ssytd2, UPLO='U':
//                        the active block stays at the top-left).  Thread i holds row i of the
//                        symmetric matrix, so p = tau A v and the rank-2 update
//                        A -= v w^T + w v^T are thread-local: per step one published pivot row,
//                        one 5-term sum and two CTA barriers.  Writes (d, e), the reflectors and
//                        their taus.
//   K2  tql_scalar_kernel  one warp per graph: the implicit-shift QL iteration (EISPACK tql2) on
//                        (d, e) only -- a ~23 000-long dependent chain of plane rotations per
//                        graph, so lane 0 runs the chain (one MUFU.RSQ and a first-order norm
//                        correction instead of sqrt and two divisions) while all lanes share the
//                        shift subtraction and the split search.  Every rotation (c, s) goes to a
//                        per-graph stream; every QL iteration is padded with identity rotations
//                        to whole 16-column blocks and described by one header word.
//   K3  tql_apply_kernel thread r holds row r of Z.  Z starts as Q = H_{n-1} ... H_1 (row r of
//                        I times the reflectors: thread-local dot + axpy against the broadcast
//                        reflector), then the rotation stream is replayed on the registers:
//                        16 statically indexed rotation stanzas per block, (c, s) pairs staged
//                        per warp by cp.async, no CTA barrier.  Epilogue through shared memory:
//                        unit columns, rank by counting, sign canonicalisation, coalesced store.
//
// Accuracy (fp32, measured on B200 against a float64 decomposition of normalised 160-node
// connectomes): eigenvalues 1e-6, eigenvectors (median) 4e-5, ||U^T U - I||_max 2e-6 -- the
// class of LAPACK ssyevd and of the Jacobi kernel after its polish phase.
#include <cstdlib>

#include "common.cuh"

namespace gsr {
int g_eigh_ql_max_iter = 30;    // gsr_set_tuning("eigh_ql_max_iter", k): QL iterations allowed per eigenvalue (EISPACK: 30)
int g_eigh_k2_lanes = 1;        // gsr_set_tuning("eigh_k2_lanes", 0|1): QL iteration with lane = graph (32 graphs per warp) for batches >= 32
                                // (default), 0 = one warp per graph
int g_eigh_k2_warps = 2;        // gsr_set_tuning("eigh_k2_warps", 1..4): warps (of 32 graphs) per CTA of the lane = graph QL kernel
int g_eigh_chunks = 8;          // gsr_set_tuning("eigh_chunks", k): the n <= 160 decomposition of a batch of >= 256 graphs runs as up to k
                                // chunks on internal streams (K2 of a chunk under K1 / K3 of the others); 0 or 1 = one launch per stage
int g_eigh_k1_regblocks = 8;   // gsr_set_tuning("eigh_k1_regblocks", 8|9|10): 16-column blocks of a row that K1 keeps in registers (the rest in
                                // shared memory); 10 = the whole row, which leaves ptxas 8 registers for everything else (round 2, first sessions)
int g_eigh_q_launch = 1;        // gsr_set_tuning("eigh_q_launch", 0|1): chunked pipeline: Q = H_{n-1} ... H_1 formed by a launch of its own right behind K1
                                // (default; K3 then only replays the rotations) / by K3 in front of its rotation replay (one launch per stage: always K3)
int g_eigh_redo = 1;            // gsr_set_tuning("eigh_redo_failed", 0|1): Jacobi pass over the graphs the QL iteration gave up on
                                // (info == -2), in the same stream (default); 0 leaves them flagged with NaN results
namespace tq {

constexpr int MAXN = 160;
constexpr int NB = MAXN / 16;            // 16-column blocks
constexpr int LDV = 164;                 // pitch of a stored reflector row: v[0..159], tau at [160]
constexpr int BIGN = 288;                // the 268-node HR target (utils/gsr_train.py:34) takes the shared-memory variants below
constexpr int LDVB = BIGN + 4;           // their reflector pitch: v[0..287], tau at [288]
__host__ __device__ inline int ldv_of(int n) { return n <= MAXN ? LDV : LDVB; }
constexpr int LDZ = 161;                 // K3 epilogue: column pitch in shared memory, odd
constexpr int BLK = 64;                  // rotations per cp.async block
constexpr int STAGES = 3;



struct Work {
    float2* stream;      // [batch][cap]      rotations (c, s) in application order (padded)
    float* V;            // [batch][n][LDV]   reflector k in row k (zeros from column k on), tau at [160]
    float* Zt;           // [batch][n][n]     K3 -> K4: column k of Z in row k (unsorted, un-normalised)
    float* d;            // [batch][n]        diagonal of T, then the unsorted eigenvalues
    float* e;            // [batch][n]        off-diagonal of T (e[n-1] = 0)
    float* scale;        // [batch]           power-of-two input scaling
    int* counts;         // [batch][4]        QL iterations, stream entries, status, -
    uint32_t* hdr;       // [batch][hdr_cap]  lo16 | hi16 << 8 of every QL iteration (16-column blocks)
    int* chunk_done;     // [64]              graphs of chunk c whose K1 has finished (chunked pipeline: K2 of the chunk waits on it)
    int cap, hdr_cap;
};

__host__ __device__ inline int stream_cap(int n) { return ((2 * n * n + 512 + BLK - 1) / BLK) * BLK + BLK; }
__host__ __device__ inline int header_cap(int n) { return 30 * n + 2; }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t carve(Work* wk, void* base, int batch, int n) {
    const int cap = stream_cap(n), hcap = header_cap(n);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_stream = take((size_t)batch * cap * sizeof(float2));
    const size_t o_v = take((size_t)batch * n * ldv_of(n) * sizeof(float));
    const size_t o_zt = take((size_t)batch * n * n * sizeof(float));
    const size_t o_d = take((size_t)batch * n * sizeof(float));
    const size_t o_e = take((size_t)batch * n * sizeof(float));
    const size_t o_sc = take((size_t)batch * sizeof(float));
    const size_t o_cnt = take((size_t)batch * 4 * sizeof(int));
    const size_t o_hdr = take((size_t)batch * hcap * sizeof(uint32_t));
    const size_t o_done = take(64 * sizeof(int));
    if (wk) {
        char* b = static_cast<char*>(base);
        wk->stream = reinterpret_cast<float2*>(b + o_stream);
        wk->V = reinterpret_cast<float*>(b + o_v);
        wk->Zt = reinterpret_cast<float*>(b + o_zt);
        wk->d = reinterpret_cast<float*>(b + o_d);
        wk->e = reinterpret_cast<float*>(b + o_e);
        wk->scale = reinterpret_cast<float*>(b + o_sc);
        wk->counts = reinterpret_cast<int*>(b + o_cnt);
        wk->hdr = reinterpret_cast<uint32_t*>(b + o_hdr);
        wk->chunk_done = reinterpret_cast<int*>(b + o_done);
        wk->cap = cap;
        wk->hdr_cap = hcap;
    }
    return off;
}


// the workspace views of graphs [g0, ...): every per-graph array advanced, the chunk counters shared
static Work offset_work(const Work& full, int g0, int n) {
    Work wk = full;
    wk.stream += (size_t)g0 * full.cap;
    wk.V += (size_t)g0 * n * ldv_of(n);
    wk.Zt += (size_t)g0 * n * n;
    wk.d += (size_t)g0 * n;
    wk.e += (size_t)g0 * n;
    wk.scale += g0;
    wk.counts += 4 * (size_t)g0;
    wk.hdr += (size_t)g0 * full.hdr_cap;
    return wk;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ------------------------------------------------------------------------------------ K1
// Registers: 5 warps per CTA and 2 CTAs per SM put 3 warps on one of the four schedulers, whose 16 K registers then
// allow 168 per thread.  With all 160 columns of the row in registers that leaves ~8 for everything else: ptxas spilled
// and -- what cost the time -- kept ONE broadcast LDS.128 in flight in the mat-vec and the rank-2 update (120 loads of
// ~33 clk per step one after the other: ~5000 clk per Householder step, FMA pipe 14 %).  RB = 16-column blocks of the
// row kept in registers; the blocks above them (columns 16 RB .. 159) live in shared memory as tail[column][thread]
// (bank = lane) and are only touched while k > 16 RB, i.e. during the first 160 - 16 RB steps.  The freed registers
// hold the broadcast operands of a whole block, loaded before the block's FMAs.  Same arithmetic in the same order:
// the variants are bit identical (tools/eigh_k1_ab.py).  Measured, 1024 graphs, one launch per stage: the decomposition
// 5.04 ms (RB = 10) -> 4.83 (9) -> 4.63 (8, four LDS.128 in flight: the default); as the chunked pipeline 3.89 -> 3.78.
template <int RB, int MINB>
__global__ void __launch_bounds__(MAXN, MINB)
tridiag_kernel(const float* __restrict__ Ain, float* __restrict__ Wout, float* __restrict__ Uout,
               int32_t* __restrict__ info, const Work wk, int n, int chunk_graphs) {
    constexpr int RC = 16 * RB;                              // columns in registers
    constexpr int TC = MAXN - RC;                            // columns in shared memory
    __shared__ __align__(16) float tile[5 * 176];            // xbuf[2][176], vbuf, wbuf, pbuf
    __shared__ __align__(16) float red[8];
    __shared__ float dd[MAXN], ee[MAXN];
    __shared__ float tail[(TC > 0 ? TC : 1) * MAXN];         // [column - RC][thread]
    float* xbuf = tile;                    // [2][176]  published pivot row (double buffered by step parity)
    float* vbuf = tile + 2 * 176;          // [176]
    float* wbuf = vbuf + 176;              // [176]
    float* pbuf = wbuf + 176;              // [176]

    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nwarps = nthr >> 5;
    const int g = blockIdx.x;
    const float* Ag = Ain + (size_t)g * n * n;
    float* mytail = tail + tid;            // column c of this thread's row: mytail[(c - RC) * MAXN]

    // ---- load: thread i gets row i of the symmetric matrix = mirror of the upper triangle:
    //      A[i][j] for j >= i (row segment), A[j][i] for j < i (coalesced over the warp)
    float a[RC];
    int bad = 0;
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        float v = 0.f;
        if (tid < n && j < n) v = __ldg(Ag + (j >= tid ? (size_t)tid * n + j : (size_t)j * n + tid));
        a[j] = v;
    }
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        if (!isfinite(a[j])) bad = 1;
        amax = fmaxf(amax, fabsf(a[j]));
    }
#pragma unroll
    for (int c = 0; c < TC; ++c) {
        const int j = RC + c;
        float v = 0.f;
        if (tid < n && j < n) v = __ldg(Ag + (j >= tid ? (size_t)tid * n + j : (size_t)j * n + tid));
        mytail[c * MAXN] = v;
        if (!isfinite(v)) bad = 1;
        amax = fmaxf(amax, fabsf(v));
    }
    if (__syncthreads_or(bad)) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int e = tid; e < n * n; e += nthr) Uout[(size_t)g * n * n + e] = qnan;
        for (int e = tid; e < n; e += nthr) Wout[(size_t)g * n + e] = qnan;
        if (tid == 0) { info[g] = -1; wk.counts[4 * g + 0] = 0; wk.counts[4 * g + 1] = 0; wk.counts[4 * g + 2] = -1; }
        if (chunk_graphs > 0) {                                        // chunked pipeline: this graph is done (see the end of the kernel)
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(wk.chunk_done + g / chunk_graphs, 1); }
        }
        return;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[w] = amax;
    __syncthreads();
    amax = 0.f;
    for (int q = 0; q < nwarps; ++q) amax = fmaxf(amax, red[q]);
    // exact power-of-two scaling into a range where the squares of the QL iteration are safe
    float sc = 1.0f;
    if (amax > 0.f && (amax < 9.5367431640625e-07f || amax > 1048576.0f)) sc = ldexpf(1.0f, -max(-120, min(120, ilogbf(amax))));
    if (sc != 1.0f) {
#pragma unroll
        for (int j = 0; j < RC; ++j) a[j] *= sc;
#pragma unroll
        for (int c = 0; c < TC; ++c) mytail[c * MAXN] *= sc;
    }
    if (tid == 0) { wk.scale[g] = sc; wk.counts[4 * g + 2] = 0; }
    __syncthreads();
    for (int e = tid; e < 5 * 176; e += nthr) tile[e] = 0.f;           // xbuf, vbuf, wbuf, pbuf
    if (tid < 8) red[tid] = 0.f;
    __syncthreads();

    // publish a row (columns 0..cnt-1, whole 16-column blocks) into xbuf[parity]
#define GSR_PUBLISH(parity, cnt)                                                                                   \
    do {                                                                                                           \
        float4* X4 = reinterpret_cast<float4*>(xbuf + (parity) * 176);                                             \
        _Pragma("unroll") for (int cb = 0; cb < RB; ++cb) {                                                        \
            if (16 * cb < (cnt)) {                                                                                 \
                _Pragma("unroll") for (int c = 0; c < 4; ++c)                                                      \
                    X4[4 * cb + c] = make_float4(a[16 * cb + 4 * c], a[16 * cb + 4 * c + 1], a[16 * cb + 4 * c + 2], a[16 * cb + 4 * c + 3]); \
            }                                                                                                      \
        }                                                                                                          \
        _Pragma("unroll") for (int cb = RB; cb < NB; ++cb) {                                                       \
            if (16 * cb < (cnt)) {                                                                                 \
                _Pragma("unroll") for (int c = 0; c < 16; ++c)                                                     \
                    xbuf[(parity) * 176 + 16 * cb + c] = mytail[(16 * (cb - RB) + c) * MAXN];                      \
            }                                                                                                      \
        }                                                                                                          \
    } while (0)
    float* Vg = wk.V + (size_t)g * n * LDV;
    if (tid == n - 1) GSR_PUBLISH((n - 1) & 1, n);
    __syncthreads();

    // ---- Householder reduction, last column first (active block = rows / columns 0..k-1)
    for (int k = n - 1; k >= 1; --k) {
        const float* xr = xbuf + (k & 1) * 176;       // row k: columns 0..k
        float xl[MAXN / 32];                           // x[lane + 32 t]
        float ss = 0.f;
#pragma unroll
        for (int t = 0; t < MAXN / 32; ++t) {
            xl[t] = xr[lane + 32 * t];                 // columns >= k of the buffer are stale: masked below
            if (lane + 32 * t < k - 1) ss = fmaf(xl[t], xl[t], ss);
        }
        ss = warp_sum(ss);                             // every warp computes the same bits
        const float alpha = xr[k - 1];
        float tau = 0.f, beta = alpha;
        if (ss != 0.f) {
            const float nrm = sqrtf(fmaf(alpha, alpha, ss));
            beta = -copysignf(nrm, alpha);
            tau = (beta - alpha) / beta;
            const float scl = 1.0f / (alpha - beta);
            // v (v[k-1] = 1, zeros above) -- every warp writes the same values
#pragma unroll
            for (int t = 0; t < MAXN / 32; ++t) {
                const int j = lane + 32 * t;
                vbuf[j] = (j < k - 1) ? xl[t] * scl : (j == k - 1 ? 1.0f : 0.0f);
            }
        }
        if (tid == 0) { dd[k] = xr[k]; ee[k - 1] = beta; }
        if (tau != 0.f) {                              // uniform
            __syncwarp();
            const float4* V4 = reinterpret_cast<const float4*>(vbuf);
            const bool act = tid < k;
            // p = tau A v: thread-local; the four broadcast loads of a block are issued before its FMAs
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            if (act) {
#pragma unroll
                for (int cb = 0; cb < RB; ++cb) {
                    if (16 * cb < k) {
                        float4 v[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) v[c] = V4[4 * cb + c];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            q0 = fmaf(a[16 * cb + 4 * c + 0], v[c].x, q0);
                            q1 = fmaf(a[16 * cb + 4 * c + 1], v[c].y, q1);
                            q2 = fmaf(a[16 * cb + 4 * c + 2], v[c].z, q2);
                            q3 = fmaf(a[16 * cb + 4 * c + 3], v[c].w, q3);
                        }
                    }
                }
#pragma unroll
                for (int cb = RB; cb < NB; ++cb) {
                    if (16 * cb < k) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 v = V4[4 * cb + c];
                            const float* t4 = mytail + (16 * (cb - RB) + 4 * c) * MAXN;
                            q0 = fmaf(t4[0], v.x, q0);
                            q1 = fmaf(t4[MAXN], v.y, q1);
                            q2 = fmaf(t4[2 * MAXN], v.z, q2);
                            q3 = fmaf(t4[3 * MAXN], v.w, q3);
                        }
                    }
                }
            }
            const float pi = tau * ((q0 + q1) + (q2 + q3));
            const float vi = act ? vbuf[tid] : 0.f;
            const float pvp = warp_sum(act ? pi * vi : 0.f);
            if (lane == 0) red[w] = pvp;
            if (act) pbuf[tid] = pi;
            // the reflector goes to global memory for K3 (whole 16-column blocks, tau at [160])
            if (w == nwarps - 1) {
                float* vrow = Vg + (size_t)k * LDV;
#pragma unroll
                for (int t = 0; t < MAXN / 32; ++t) vrow[lane + 32 * t] = vbuf[lane + 32 * t];
                if (lane == 0) vrow[MAXN] = tau;
            }
            __syncthreads();                                                   // (A)
            const float4 r0 = *reinterpret_cast<const float4*>(red), r1 = *reinterpret_cast<const float4*>(red + 4);
            const float pv = ((r0.x + r0.y) + (r0.z + r0.w)) + ((r1.x + r1.y) + (r1.z + r1.w));
            const float coef = 0.5f * tau * pv;
            // w = p - coef v -- every warp writes the same values
#pragma unroll
            for (int t = 0; t < MAXN / 32; ++t) {
                const int j = lane + 32 * t;
                wbuf[j] = (j < k) ? fmaf(-coef, vbuf[j], pbuf[j]) : 0.0f;
            }
            __syncwarp();
            if (act) {
                const float wi = wbuf[tid];
                const float4* W4 = reinterpret_cast<const float4*>(wbuf);
#pragma unroll
                for (int cb = 0; cb < RB; ++cb) {
                    if (16 * cb < k) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                          // 8 columns: two v and two w quads in flight
                            const float4 v0 = V4[4 * cb + 2 * h], v1 = V4[4 * cb + 2 * h + 1];
                            const float4 w0 = W4[4 * cb + 2 * h], w1 = W4[4 * cb + 2 * h + 1];
                            float* ar = a + 16 * cb + 8 * h;
                            ar[0] = fmaf(-vi, w0.x, fmaf(-wi, v0.x, ar[0]));
                            ar[1] = fmaf(-vi, w0.y, fmaf(-wi, v0.y, ar[1]));
                            ar[2] = fmaf(-vi, w0.z, fmaf(-wi, v0.z, ar[2]));
                            ar[3] = fmaf(-vi, w0.w, fmaf(-wi, v0.w, ar[3]));
                            ar[4] = fmaf(-vi, w1.x, fmaf(-wi, v1.x, ar[4]));
                            ar[5] = fmaf(-vi, w1.y, fmaf(-wi, v1.y, ar[5]));
                            ar[6] = fmaf(-vi, w1.z, fmaf(-wi, v1.z, ar[6]));
                            ar[7] = fmaf(-vi, w1.w, fmaf(-wi, v1.w, ar[7]));
                        }
                    }
                }
#pragma unroll
                for (int cb = RB; cb < NB; ++cb) {
                    if (16 * cb < k) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 v = V4[4 * cb + c], ww = W4[4 * cb + c];
                            float* t4 = mytail + (16 * (cb - RB) + 4 * c) * MAXN;
                            t4[0] = fmaf(-vi, ww.x, fmaf(-wi, v.x, t4[0]));
                            t4[MAXN] = fmaf(-vi, ww.y, fmaf(-wi, v.y, t4[MAXN]));
                            t4[2 * MAXN] = fmaf(-vi, ww.z, fmaf(-wi, v.z, t4[2 * MAXN]));
                            t4[3 * MAXN] = fmaf(-vi, ww.w, fmaf(-wi, v.w, t4[3 * MAXN]));
                        }
                    }
                }
            }
        } else if (w == nwarps - 1 && lane == 0) {
            Vg[(size_t)k * LDV + MAXN] = 0.f;                                  // H = I (slarfg): K3 skips the row
        }
        // the next pivot row (row k-1, columns 0..k-1) is final now
        if (tid == k - 1) GSR_PUBLISH((k - 1) & 1, k);
        __syncthreads();                                                       // (B)
    }
#undef GSR_PUBLISH
    if (tid == 0) {
        dd[0] = (n == 1) ? a[0] : xbuf[0];              // row 0 was published last (buffer 0, column 0)
        ee[n - 1] = 0.f;
        Vg[MAXN] = 0.f;                                 // no reflector in row 0
    }
    __syncthreads();
    for (int j = tid; j < n; j += nthr) {
        wk.d[(size_t)g * n + j] = dd[j];
        wk.e[(size_t)g * n + j] = (j < n - 1) ? ee[j] : 0.f;
    }
    if (chunk_graphs > 0) {
        // chunked pipeline: everything this CTA wrote (d, e, reflectors, counts) is visible device-wide before the count of
        // finished graphs of its chunk goes up; the K2 CTA of the chunk -- resident since the fork -- polls that count
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(wk.chunk_done + g / chunk_graphs, 1); }
    }
}

// ------------------------------------------------------------------------------------ K2
// EISPACK tql2 on (d, e); rotations are recorded instead of applied.
constexpr int K2_WARPS = 4;

__global__ void __launch_bounds__(32 * K2_WARPS)
tql_scalar_kernel(const Work wk, int n, int batch, int max_iter) {
    __shared__ float sd[K2_WARPS][BIGN + 16], se[K2_WARPS][BIGN + 16];      // [8] pad | n values | pad: index -1 is readable
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int g = blockIdx.x * K2_WARPS + wl;
    if (g >= batch) return;
    int* cnt = wk.counts + 4 * g;
    if (cnt[2] == -1) return;                       // non-finite input (K1 wrote the outputs)
    float* d = sd[wl] + 8;
    float* e = se[wl] + 8;
    for (int j = lane - 8; j < n + 8; j += 32) {
        d[j] = (j >= 0 && j < n) ? wk.d[(size_t)g * n + j] : 0.f;
        e[j] = (j >= 0 && j < n - 1) ? wk.e[(size_t)g * n + j] : 0.f;
    }
    __syncwarp();
    float2* out = wk.stream + (size_t)g * wk.cap;
    uint32_t* hdr = wk.hdr + (size_t)g * wk.hdr_cap;
    const float2 ident = make_float2(1.0f, 0.0f);
    float f = 0.f, tst1 = 0.f;
    int pos = 0, nit = 0, status = 0;
    for (int l = 0; l < n && status == 0; ++l) {
        tst1 = fmaxf(tst1, fabsf(d[l]) + fabsf(e[l]));
        // smallest m >= l with a negligible e[m] (e[n-1] = 0 always qualifies)
        int m = l;
        for (int base = l;; base += 32) {
            const int idx = base + lane;
            const bool neg = (idx >= n - 1) || (tst1 + fabsf(e[idx]) == tst1);
            const unsigned b = __ballot_sync(0xffffffffu, neg);
            if (b) { m = base + __ffs(b) - 1; break; }
        }
        if (m > l) {
            // the rotations i = m-1 .. l of one QL iteration are padded with identities to the
            // 16-column blocks [lo, hi): K3 replays whole blocks with static register indices
            const int lo = l & ~15, hi = (m + 15) & ~15;
            int j = 0;
            while (true) {
                if (++j > max_iter) { status = 2; break; }
                if (pos + (hi - lo) > wk.cap - BLK || nit >= wk.hdr_cap) { status = 3; break; }
                // Wilkinson shift
                const float g0 = d[l], el = e[l];
                const float p0 = (d[l + 1] - g0) / (2.0f * el);
                const float r0 = sqrtf(fmaf(p0, p0, 1.0f));
                const float den = p0 + copysignf(r0, p0);
                const float dl = el / den, dl1 = el * den;
                const float h = g0 - dl;
                __syncwarp();
                if (lane == 0) { d[l] = dl; d[l + 1] = dl1; }
                for (int q = l + 2 + lane; q < n; q += 32) d[q] -= h;
                f += h;
                float2* o = out + pos;
                for (int q = lane; q < hi - m; q += 32) o[q] = ident;                       // rotation indices hi-1 .. m
                for (int q = lane; q < l - lo; q += 32) o[(hi - l) + q] = ident;            // rotation indices l-1 .. lo
                __syncwarp();
                if (lane == 0) {
                    hdr[nit] = (uint32_t)(lo >> 4) | ((uint32_t)(hi >> 4) << 8);
                    o += hi - m;
                    float p = d[m], c = 1.0f, s = 0.0f;
                    float s_prev = 0.0f, c_prev = 1.0f, c_prev2 = 1.0f;    // s of the previous, c of the previous two rotations
                    const float el1 = e[l + 1];
                    float ei = e[m - 1], di = d[m - 1];
#pragma unroll 4
                    for (int q = m - 1; q >= l; --q) {
                        const float ein = e[q - 1], din = d[q - 1];             // next rotation's operands (index -1 is padding)
                        const float gg = c * ei, hh = c * p;
                        const float x = fmaf(p, p, ei * ei);
                        // (c, s) = (p, e_i) / r with one MUFU: rsqrt's relative error delta shows as
                        // 1 - x rinv^2 = 2 delta, and one first-order step removes it (unit norm to ~1e-7,
                        // unbiased; a biased norm would drift the eigenvalues and the orthogonality of Z over
                        // ~300 rotations per column).  The running p of the next rotation is linear in (c, s):
                        // the step is applied to the product, so p does not wait for the corrected c, s --
                        // the loop-carried chain is x, rsqrt, rinv^2, hd, one FFMA.
                        float rinv;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(x));
                        const float hd = fmaf(-0.5f * x, rinv * rinv, 0.5f);
                        const float c0 = p * rinv, s0 = ei * rinv, xr = x * rinv;
                        const float t = fmaf(c0, di, -s0 * gg);
                        float cn = fmaf(c0, hd, c0), sn = fmaf(s0, hd, s0), r = fmaf(xr, hd, xr), pn = fmaf(t, hd, t);
                        if (x < 1.17549435e-38f) { r = 0.f; cn = 1.0f; sn = 0.0f; pn = di; }  // p = e_i = 0 (to fp32 range): identity
                        e[q + 1] = s * r;
                        s_prev = s; c_prev2 = c_prev; c_prev = c;
                        s = sn; c = cn;
                        p = pn;
                        d[q + 1] = hh + s * fmaf(c, gg, s * di);
                        *o++ = make_float2(c, s);
                        ei = ein; di = din;
                    }
                    // tql2's closing step needs s of the previous and c of the one-before-previous rotation
                    // (carried in registers: reading them back from the stream just written costs a global
                    // round trip per iteration)
                    const float s2 = s_prev, c3 = c_prev2;
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                }
                pos += hi - lo;
                ++nit;
                __syncwarp();
                if (tst1 + fabsf(e[l]) <= tst1) break;
            }
        }
        __syncwarp();
        if (lane == 0) d[l] += f;
        __syncwarp();
    }
    for (int j = lane; j < n; j += 32) wk.d[(size_t)g * n + j] = d[j];
    if (lane == 0) { cnt[0] = nit; cnt[1] = pos; cnt[2] = status; }
}

// K2, lane = graph.  The kernel above keeps ONE lane of a warp busy per graph: with 1024 graphs resident it is bound by
// instruction issue (7 warps per SM x ~40 instructions per rotation against a ~45-cycle dependent chain), i.e. 1/32 of the
// issued lanes do work.  Here a warp runs the QL iterations of 32 graphs in lockstep, lane = graph: 32 warps for 1024
// graphs, every one alone on its scheduler, so the time is the chain latency and the kernel leaves the SMs free for whatever
// runs next to it.  Per-lane state (l, m, the tridiagonal in shared memory at [index][lane]: bank = lane whatever the index);
// the lanes meet once per QL iteration ("trip"); the rotation loop of a trip runs to the longest lane's m - l.
// Differences to the one-lane kernel, all inside the freedom of the algorithm:
//   * the split search is not repeated before every iteration: the smallest index whose new off-diagonal is negligible is
//     tracked while the rotations of an iteration produce it, and serves the next iteration (also after l advances by one;
//     a full scan runs only at the start and after two eigenvalues deflate in a row).  It is judged with the tst1 of the
//     iteration that produced it -- never larger than the current one, so a split may be found an iteration late, never
//     wrongly;
//   * the shift h is subtracted from d[l+2 .. m] as the rotation loop reads them (same values), from d[m+1 ..] in a loop.
// Stream / header format and all arithmetic of a rotation are those of tql_scalar_kernel.
constexpr int K2L_WARPS = 4;            // at most one warp per scheduler; the launcher picks g_eigh_k2_warps <= 4 (they share the SM's LSU)
__global__ void __launch_bounds__(32 * K2L_WARPS)
tql_scalar_lanes_kernel(const Work wk, int n, int g_first, int batch, int max_iter, const int* wait_count, int wait_target) {
    extern __shared__ __align__(16) float sm_all[];
    // graphs [g_first, batch): warp (blockIdx.x, wl) takes 32 of them.  Chunked pipeline: this CTA has been resident since the
    // fork (it owns its SM: the launch asks for all of the shared memory) and polls the number of graphs of its chunk whose
    // K1 has finished; K1 publishes with __threadfence + atomicAdd, the acquire load + barrier order every thread's reads.
    if (wait_count) {
        if (threadIdx.x == 0) {
            atomicAdd(wk.chunk_done + 32, 1);                    // resident: the gate in front of K1 counts these
            for (unsigned spin = 0;; ++spin) {
                int v;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(wait_count) : "memory");
                if (v >= wait_target) break;
                if (spin > (1u << 22)) { printf("eigh K2: poll timeout, block %d count %d target %d\n", (int)blockIdx.x, v, wait_target); __trap(); }   // a protocol bug must fault, not hang the GPU
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int np = n + 16;                                   // 8 pad | n | 8 pad per array: index -1 and n .. n+3 are readable
    float* sm = sm_all + (size_t)wl * 2 * np * 32;
    const int g0 = g_first + (blockIdx.x * (blockDim.x >> 5) + wl) * 32, g = g0 + lane;
    if (g0 >= batch) return;
    float* d = sm + 8 * 32 + lane;                           // d[j] at d[32 * j]
    float* e = sm + (np + 8) * 32 + lane;
    for (int gg = 0; gg < 32; ++gg) {                        // coalesced over j per graph
        const int gi = g0 + gg;
        const bool ok = gi < batch && wk.counts[4 * gi + 2] != -1;
        for (int j = lane - 8; j < n + 8; j += 32) {
            float dv = 0.f, ev = 0.f;
            if (ok && j >= 0 && j < n) { dv = wk.d[(size_t)gi * n + j]; if (j < n - 1) ev = wk.e[(size_t)gi * n + j]; }
            sm[(8 + j) * 32 + gg] = dv;
            sm[(np + 8 + j) * 32 + gg] = ev;
        }
    }
    __syncwarp();
    const bool live = g < batch && wk.counts[4 * min(g, batch - 1) + 2] != -1;      // -1: non-finite input (K1 wrote the outputs)
    float2* out = wk.stream + (size_t)min(g, batch - 1) * wk.cap;
    uint32_t* hdr = wk.hdr + (size_t)min(g, batch - 1) * wk.hdr_cap;
    const float2 ident = make_float2(1.0f, 0.0f);
    int status = live ? 0 : 1;                               // 1: lane without a graph
    int l = 0, m = 0, jc = 0, pos = 0, nit = 0;
    float f = 0.f, tst1 = 0.f;
    auto negl = [&](float v) { return tst1 + fabsf(v) == tst1; };
    auto scan = [&](int from) {                              // smallest m >= from with a negligible e[m]; e[n-1] = 0 ends it
        int mm = from;
        for (;;) {
            const float v0 = e[32 * mm], v1 = e[32 * (mm + 1)], v2 = e[32 * (mm + 2)], v3 = e[32 * (mm + 3)];
            if (negl(v0)) break;
            if (negl(v1)) { mm += 1; break; }
            if (negl(v2)) { mm += 2; break; }
            if (negl(v3)) { mm += 3; break; }
            mm += 4;
        }
        return min(mm, n - 1);
    };
    auto deflate = [&]() {                                   // eigenvalues that need no (further) iteration: m == l
        while (m == l) {
            d[32 * l] += f;
            jc = 0;
            if (++l >= n) break;
            tst1 = fmaxf(tst1, fabsf(d[32 * l]) + fabsf(e[32 * l]));
            m = scan(l);
        }
    };
    if (status == 0) {
        tst1 = fabsf(d[0]) + fabsf(e[0]);
        m = scan(0);
        deflate();
    }
    for (;;) {
        const bool run = status == 0 && l < n;
        if (!__any_sync(0xffffffffu, run)) break;
        if (!run) continue;
        const int lo = l & ~15, hi = (m + 15) & ~15;
        if (++jc > max_iter) { status = 2; continue; }
        if (pos + (hi - lo) > wk.cap - BLK || nit >= wk.hdr_cap) { status = 3; continue; }
        // Wilkinson shift
        const float gl = d[32 * l], el = e[32 * l];
        const float p0 = (d[32 * (l + 1)] - gl) / (2.0f * el);
        const float r0 = sqrtf(fmaf(p0, p0, 1.0f));
        const float den = p0 + copysignf(r0, p0);
        const float dl = el / den, dl1 = el * den;
        const float h = gl - dl;
        d[32 * l] = dl; d[32 * (l + 1)] = dl1;
        for (int q = max(m + 1, l + 2); q < n; ++q) d[32 * q] -= h;          // below a split (rare); d[l+2 .. m] get it on the fly
        f += h;
        // The entries of an iteration (identities for the rotation indices hi-1 .. m, the rotations m-1 .. l, identities for
        // l-1 .. lo) go out as 16-byte PAIRS: position P of an entry inside the iteration is hi - 1 - (rotation index), the
        // iteration starts at a multiple of 16 entries, so (even P, odd P) is one aligned float4.  A lane-strided store costs
        // the LSU one sector request per lane whatever its width; pairs halve what was 32 requests per rotation and warp.
        float4* o4 = reinterpret_cast<float4*>(out + pos);                   // pair k = positions 2k, 2k + 1
        hdr[nit] = (uint32_t)(lo >> 4) | ((uint32_t)(hi >> 4) << 8);
        const float4 ident2 = make_float4(1.0f, 0.0f, 1.0f, 0.0f);
        for (int k = 0; k < (hi - m) >> 1; ++k) o4[k] = ident2;              // whole pairs of leading identities
        int P = hi - m;                                                      // position of the next rotation
        float p = d[32 * m] - (m >= l + 2 ? h : 0.f), c = 1.0f, s = 0.0f;
        float s_prev = 0.0f, c_prev = 1.0f, c_prev2 = 1.0f;
        const float el1 = e[32 * (l + 1)];
        float ei = e[32 * (m - 1)], di = d[32 * (m - 1)] - (m - 1 >= l + 2 ? h : 0.f);
        int mneg = m;                                                        // e[m] becomes s r = 0
        const int lsub = l + 3;                                              // rotation q reads d[q - 1]: shifted iff q >= l + 3
        auto rotate = [&](int q) {                                           // rotation index q: updates (c, s), p, d / e[q + 1]
            const float ein = e[32 * (q - 1)];                               // next rotation's operands (index -1 is padding)
            const float din = d[32 * (q - 1)] - (q >= lsub ? h : 0.f);
            const float gg = c * ei, hh = c * p;
            const float x = fmaf(p, p, ei * ei);
            float rinv;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(x));
            const float hd = fmaf(-0.5f * x, rinv * rinv, 0.5f);             // see tql_scalar_kernel: first-order unit-norm correction
            const float c0 = p * rinv, s0 = ei * rinv, xr = x * rinv;
            const float t = fmaf(c0, di, -s0 * gg);
            float cn = fmaf(c0, hd, c0), sn = fmaf(s0, hd, s0), r = fmaf(xr, hd, xr), pn = fmaf(t, hd, t);
            if (__builtin_expect(x < 1.17549435e-38f, 0)) { r = 0.f; cn = 1.0f; sn = 0.0f; pn = di; }
            const float enew = s * r;
            e[32 * (q + 1)] = enew;
            if (negl(enew)) mneg = q + 1;                                    // q descends: the last hit is the smallest index
            s_prev = s; c_prev2 = c_prev; c_prev = c;
            s = sn; c = cn;
            p = pn;
            d[32 * (q + 1)] = hh + s * fmaf(c, gg, s * di);
            ei = ein; di = din;
        };
        int q = m - 1;
        if (P & 1) {                                                         // an odd first position pairs with the identity before it
            rotate(q);
            o4[P >> 1] = make_float4(1.0f, 0.0f, c, s);
            ++P; --q;
        }
        for (; q > l; q -= 2) {                                              // two rotations, one 16-byte store
            rotate(q);
            const float ca = c, sa = s;
            rotate(q - 1);
            o4[P >> 1] = make_float4(ca, sa, c, s);
            P += 2;
        }
        if (q == l) {                                                        // a last rotation at an even position pairs with the first
            rotate(q);                                                       // trailing identity (l is odd then, so l > lo)
            o4[P >> 1] = make_float4(c, s, 1.0f, 0.0f);
            P += 2;
        }
        p = -s * s_prev * c_prev2 * el1 * e[32 * l] / dl1;
        e[32 * l] = s * p;
        d[32 * l] = c * p;
        for (int k = P >> 1; k < (hi - lo) >> 1; ++k) o4[k] = ident2;        // whole pairs of trailing identities
        pos += hi - lo;
        ++nit;
        if (tst1 + fabsf(e[32 * l]) <= tst1) {                               // eigenvalue l converged
            d[32 * l] += f;
            jc = 0;
            if (++l < n) {
                tst1 = fmaxf(tst1, fabsf(d[32 * l]) + fabsf(e[32 * l]));
                m = negl(e[32 * l]) ? l : max(mneg, l);
                deflate();
            }
        } else {
            m = mneg;
        }
    }
    __syncwarp();
    for (int gg = 0; gg < 32; ++gg) {
        const int gi = g0 + gg;
        if (gi < batch && wk.counts[4 * gi + 2] != -1)
            for (int j = lane; j < n; j += 32) wk.d[(size_t)gi * n + j] = sm[(8 + j) * 32 + gg];
    }
    if (live) { int* cnt = wk.counts + 4 * g; cnt[0] = nit; cnt[1] = pos; cnt[2] = status; }
}

// Gate in front of K1 (chunked pipeline): one thread waits -- for at most ~60 us -- until the K2 CTAs have checked in as
// resident.  They are launched first, but on other streams: without the gate K1's 1024 CTAs can reach the SMs a few
// microseconds before the last K2 launch and then that chunk's K2 CTAs (each needs an empty SM) are placed only when K1
// drains (measured: 3.9 -> 5.5 ms).  The wait is bounded: if the SMs are busy with something else K1 simply starts.
__global__ void k2_resident_gate_kernel(const int* resident, int expected) {
    for (int spin = 0; spin < 600; ++spin) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(resident) : "memory");
        if (v >= expected) break;
        __nanosleep(100);
    }
}

// ------------------------------------------------------------------------------------ K3

constexpr int RING_FLOATS = STAGES * 176;       // per warp: 3 reflector rows (164 floats) or 3 x 64 (c, s) pairs (128 floats)
constexpr int K3_WARPS = 4;                     // warp units per CTA; 2 CTAs per SM = 2 warps per scheduler at <= 255 registers

// Row `tid` of Q = H_{n-1} ... H_1 (e_r^T H_{n-1} ... H_1: a thread-local dot + axpy per reflector) into z; the reflector
// rows n-1 .. 1 stream through the warp's 3-deep cp.async ring (41 x 16 B each).
__device__ __forceinline__ void form_q_row(float (&z)[MAXN + 1], const Work& wk, int g, int n, int tid, int lane, float* ring) {
#pragma unroll
    for (int j = 0; j <= MAXN; ++j) z[j] = (j == tid && tid < n) ? 1.0f : 0.0f;
    const float* Vg = wk.V + (size_t)g * n * LDV;
    // reflector rows n-1 .. 1 stream through a 3-deep per-warp ring (41 x 16 B each)
    auto issue = [&](int k) {
        if (k >= 1) {
            const float* src = Vg + (size_t)k * LDV;
            float* dst = ring + (k % STAGES) * 176;
            cp_async16(dst + 4 * lane, src + 4 * lane);
            if (lane < LDV / 4 - 32) cp_async16(dst + 4 * (lane + 32), src + 4 * (lane + 32));
        }
        cp_async_commit();
    };
    issue(n - 1);
    issue(n - 2);
    for (int k = n - 1; k >= 1; --k) {
        cp_async_wait<1>();
        __syncwarp();                                              // row k landed; every lane is done with row k + 1
        issue(k - 2);                                              // stage (k + 1) % 3: consumed one step ago
        const float* vr = ring + (k % STAGES) * 176;
        const float tau = vr[MAXN];
        if (tau != 0.f) {                                          // uniform
            const float4* V4 = reinterpret_cast<const float4*>(vr);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int cb = 0; cb < NB; ++cb) {
                if (16 * cb < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = V4[4 * cb + c];
                        q0 = fmaf(z[16 * cb + 4 * c + 0], v.x, q0);
                        q1 = fmaf(z[16 * cb + 4 * c + 1], v.y, q1);
                        q2 = fmaf(z[16 * cb + 4 * c + 2], v.z, q2);
                        q3 = fmaf(z[16 * cb + 4 * c + 3], v.w, q3);
                    }
                }
            }
            const float s = -tau * ((q0 + q1) + (q2 + q3));
#pragma unroll
            for (int cb = 0; cb < NB; ++cb) {
                if (16 * cb < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = V4[4 * cb + c];
                        z[16 * cb + 4 * c + 0] = fmaf(s, v.x, z[16 * cb + 4 * c + 0]);
                        z[16 * cb + 4 * c + 1] = fmaf(s, v.y, z[16 * cb + 4 * c + 1]);
                        z[16 * cb + 4 * c + 2] = fmaf(s, v.z, z[16 * cb + 4 * c + 2]);
                        z[16 * cb + 4 * c + 3] = fmaf(s, v.w, z[16 * cb + 4 * c + 3]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
}

// COLUMN `col` of Q = H_{n-1} ... H_1 into z (z[j] = Q[j][col]): H_1 is applied first, and H_k -- whose vector lives in
// coordinates [0, k) -- leaves e_col alone for every k <= col, so a warp that owns columns c0 .. c0 + 31 starts at
// k = c0 + 1 (lanes whose column is still >= k see a zero dot product: no predicate needed).  4/3 n^3 flops over the graph
// instead of the 2 n^3 of the row form above (there every row is dense from the first reflector on); same inner loops, the
// reflector rows stream through the ring in ascending order.
__device__ __forceinline__ void form_q_col(float (&z)[MAXN + 1], const Work& wk, int g, int n, int col, int c0, int lane, float* ring) {
#pragma unroll
    for (int j = 0; j <= MAXN; ++j) z[j] = (j == col && col < n) ? 1.0f : 0.0f;
    const float* Vg = wk.V + (size_t)g * n * LDV;
    auto issue = [&](int k) {
        if (k < n) {
            const float* src = Vg + (size_t)k * LDV;
            float* dst = ring + (k % STAGES) * 176;
            cp_async16(dst + 4 * lane, src + 4 * lane);
            if (lane < LDV / 4 - 32) cp_async16(dst + 4 * (lane + 32), src + 4 * (lane + 32));
        }
        cp_async_commit();
    };
    const int k_first = c0 + 1;
    issue(k_first);
    issue(k_first + 1);
    for (int k = k_first; k < n; ++k) {
        cp_async_wait<1>();
        __syncwarp();                                              // row k landed; every lane is done with row k - 1
        issue(k + 2);                                              // stage (k - 1) % 3: consumed one step ago
        const float* vr = ring + (k % STAGES) * 176;
        const float tau = vr[MAXN];
        if (tau != 0.f) {                                          // uniform
            const float4* V4 = reinterpret_cast<const float4*>(vr);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int cb = 0; cb < NB; ++cb) {
                if (16 * cb < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = V4[4 * cb + c];
                        q0 = fmaf(z[16 * cb + 4 * c + 0], v.x, q0);
                        q1 = fmaf(z[16 * cb + 4 * c + 1], v.y, q1);
                        q2 = fmaf(z[16 * cb + 4 * c + 2], v.z, q2);
                        q3 = fmaf(z[16 * cb + 4 * c + 3], v.w, q3);
                    }
                }
            }
            const float s = -tau * ((q0 + q1) + (q2 + q3));
#pragma unroll
            for (int cb = 0; cb < NB; ++cb) {
                if (16 * cb < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = V4[4 * cb + c];
                        z[16 * cb + 4 * c + 0] = fmaf(s, v.x, z[16 * cb + 4 * c + 0]);
                        z[16 * cb + 4 * c + 1] = fmaf(s, v.y, z[16 * cb + 4 * c + 1]);
                        z[16 * cb + 4 * c + 2] = fmaf(s, v.z, z[16 * cb + 4 * c + 2]);
                        z[16 * cb + 4 * c + 3] = fmaf(s, v.w, z[16 * cb + 4 * c + 3]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
}

constexpr int QT_PITCH = 33;                         // Kq's per-warp transpose tile: [MAXN rows][33]
constexpr int KQ_SMEM_FLOATS = K3_WARPS * (RING_FLOATS + MAXN * QT_PITCH);

// Kq (chunked pipeline, eigh_q_launch = 1): Q of every graph as its own launch right behind K1, written to Zt[column][row];
// K3 then only replays the rotations.  It fills the SMs between the end of K1 and the end of the first chunks' rotation
// chains (K2), which would otherwise idle, and takes the Q formation off the critical path K1 -> K2 -> K3 of the last chunk.
__global__ void __launch_bounds__(32 * K3_WARPS, 2)
form_q_kernel(const Work wk, int n, int units_per_graph, int total_units) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // a CTA takes ONE block of 32 columns for four consecutive graphs: its warps then run the same number of reflectors
    // (columns c0 .. c0 + 31 start at k = c0 + 1), so no warp idles inside a resident CTA while the longest one finishes
    const int cblk = blockIdx.x % units_per_graph, g = (blockIdx.x / units_per_graph) * K3_WARPS + w;
    if (g * units_per_graph >= total_units) return;
    const int c0 = cblk * 32;                                                                 // the warp's 32 columns of Q
    if (wk.counts[4 * g + 2] == -1) return;                            // non-finite input: K1 wrote the outputs (K2 may be updating the status of the others)
    float z[MAXN + 1];
    form_q_col(z, wk, g, n, c0 + lane, c0, lane, sm + w * RING_FLOATS);
    // Zt[column][row]: the warp holds 32 columns; through its shared-memory tile every column leaves as coalesced runs of rows
    float* tile = sm + K3_WARPS * RING_FLOATS + w * (MAXN * QT_PITCH);
#pragma unroll
    for (int j = 0; j < MAXN; ++j) tile[j * QT_PITCH + lane] = z[j];
    __syncwarp();
    float* Zg = wk.Zt + (size_t)g * n * n;
    for (int col = 0; col < 32 && c0 + col < n; ++col) {
        float* dst = Zg + (size_t)(c0 + col) * n;
#pragma unroll
        for (int rb = 0; rb < MAXN / 32; ++rb) {
            const int row = 32 * rb + lane;
            if (row < n) dst[row] = tile[row * QT_PITCH + col];
        }
    }
}

// One WARP per (graph, block of 32 rows of Z): rows are independent under column operations, so
// the unit of work is a warp and nothing synchronises across warps.  Output: Zt[g][col][row]
// (unsorted, un-normalised; coalesced: a warp stores 32 consecutive rows of one column).
__global__ void __launch_bounds__(32 * K3_WARPS, 2)
tql_apply_kernel(const Work wk, int n, int units_per_graph, int total_units, int q_from_zt) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int unit = blockIdx.x * K3_WARPS + w;
    if (unit >= total_units) return;
    const int g = unit / units_per_graph, tid = (unit - g * units_per_graph) * 32 + lane;     // tid = row of Z
    float* ring = sm + w * RING_FLOATS;
    const int* cnt = wk.counts + 4 * g;
    const int nit = cnt[0], nent = cnt[1], status = cnt[2];
    if (status != 0) return;                                           // K1 / K4 write the error outputs

    // ---- Z = Q = H_{n-1} ... H_1, row `tid` of it: e_r^T H_{n-1} ... H_1, thread-local
    float z[MAXN + 1];
    if (q_from_zt) {
        // Kq formed Q (eigh_q_launch): row `tid` of it comes back from Zt[column][row], coalesced over the warp
        const float* Zg = wk.Zt + (size_t)g * n * n;
#pragma unroll
        for (int j = 0; j <= MAXN; ++j) z[j] = (j < n && tid < n) ? Zg[(size_t)j * n + tid] : 0.0f;
    } else {
        form_q_row(z, wk, g, n, tid, lane, ring);
    }

    // ---- replay the rotation stream: iteration = blocks hb-1 .. lb of 16 rotations, rotation
    //      index i acts on columns (i, i + 1), i descending
    {
        const float2* src = wk.stream + (size_t)g * wk.cap;
        const uint32_t* hdr = wk.hdr + (size_t)g * wk.hdr_cap;
        float2* cring = reinterpret_cast<float2*>(ring);               // [STAGES][BLK]
        const int nblk = (nent + BLK - 1) / BLK;
        auto issue = [&](int b) {
            if (b < nblk) cp_async16(cring + (b % STAGES) * BLK + 2 * lane, src + (size_t)b * BLK + 2 * lane);
            cp_async_commit();
        };
        issue(0);
        issue(1);
        // `slot`: where the next 16 rotations sit in the ring (float2 index, = ((pos / BLK) % STAGES) * BLK + pos % BLK for the
        // running position pos), kept incrementally; `nextblk`: the next 64-rotation block to request.  An iteration ENTERS the
        // statically indexed stanzas at block hb - 1 through a jump table and falls through down to block lb: one compare per
        // executed stanza, nothing for the others (the ten-way if chain of the first version, with a division by 3 per stanza
        // for the ring slot, was ~45 % of this kernel's stall samples: profiles/r2c_ncu_summary_eigh.txt).
        int slot = 0, nextblk = 2;
        uint32_t h = nit > 0 ? __ldg(hdr) : 0u;
#define GSR_STANZA(b)                                                                                              \
        {                                                                                                          \
            if ((slot & (BLK - 1)) == 0) {                     /* entering a new 64-rotation block */              \
                cp_async_wait<1>();                                                                                \
                __syncwarp();                                                                                      \
                issue(nextblk);                                                                                    \
                ++nextblk;                                                                                         \
            }                                                                                                      \
            const float4* cs4 = reinterpret_cast<const float4*>(cring + slot);                                     \
            _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                                        \
                const float4 q = cs4[u];                       /* two rotations: (c, s) of i, (c, s) of i - 1 */   \
                constexpr int i0 = 16 * (b) + 15;                                                                  \
                const int i = i0 - 2 * u;                                                                          \
                /* the running column t enters each product last, so the dependent chain is one FFMA per */       \
                /* rotation (the products with z[i], z[i-1] are off it) */                                         \
                float t = z[i + 1];                                                                                \
                z[i + 1] = fmaf(q.x, t, q.y * z[i]);                                                               \
                z[i] = fmaf(-q.y, t, q.x * z[i]);                                                                  \
                t = z[i];                                                                                          \
                z[i] = fmaf(q.z, t, q.w * z[i - 1]);                                                               \
                z[i - 1] = fmaf(-q.w, t, q.z * z[i - 1]);                                                          \
            }                                                                                                      \
            slot += 16;                                                                                            \
            if (slot == STAGES * BLK) slot = 0;                                                                    \
        }
        for (int it = 0; it < nit; ++it) {
            const uint32_t hn = (it + 1 < nit) ? __ldg(hdr + it + 1) : 0u;
            const int lb = (int)(h & 0xffu), hb = (int)((h >> 8) & 0xffu);
            static_assert(NB == 10, "the stanza switch below lists blocks 9 .. 0");
            if (hb > lb) {                                                 // uniform
                switch (min(hb, NB) - 1) {
                    case 9: GSR_STANZA(9) if (lb == 9) break; [[fallthrough]];
                    case 8: GSR_STANZA(8) if (lb == 8) break; [[fallthrough]];
                    case 7: GSR_STANZA(7) if (lb == 7) break; [[fallthrough]];
                    case 6: GSR_STANZA(6) if (lb == 6) break; [[fallthrough]];
                    case 5: GSR_STANZA(5) if (lb == 5) break; [[fallthrough]];
                    case 4: GSR_STANZA(4) if (lb == 4) break; [[fallthrough]];
                    case 3: GSR_STANZA(3) if (lb == 3) break; [[fallthrough]];
                    case 2: GSR_STANZA(2) if (lb == 2) break; [[fallthrough]];
                    case 1: GSR_STANZA(1) if (lb == 1) break; [[fallthrough]];
                    default: GSR_STANZA(0)
                }
            }
            h = hn;
        }
#undef GSR_STANZA
        cp_async_wait<0>();
    }
    if (tid < n) {
        float* Zg = wk.Zt + (size_t)g * n * n;
#pragma unroll
        for (int j = 0; j < MAXN; ++j) {
            if (j < n) Zg[(size_t)j * n + tid] = z[j];
        }
    }
}

// ------------------------------------------------------------------------------------ K4
// One CTA per graph: unit columns, ascending order by counting, sign rule, coalesced store.
__global__ void __launch_bounds__(256)
tql_finish_kernel(const Work wk, float* __restrict__ Wout, float* __restrict__ Uout, int32_t* __restrict__ info,
                  int n, int canonical) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm;                                                    // column k of Z at zs + k * LDZ
    float* lam = sm + (size_t)n * LDZ;                                 // [n]
    float* sgn = lam + n;                                              // [n]
    int* inv = reinterpret_cast<int*>(sgn + n);                        // [n]
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nwarps = nthr >> 5;
    const int g = blockIdx.x;
    const int* cnt = wk.counts + 4 * g;
    const int nit = cnt[0], status = cnt[2];
    if (status == -1) return;                                          // K1 wrote NaNs and info
    float* Wg = Wout + (size_t)g * n;
    float* Ug = Uout + (size_t)g * n * n;
    if (status != 0) {                                                 // QL did not converge / stream overflow
        const float qnan = __int_as_float(0x7fc00000);
        for (int e = tid; e < n * n; e += nthr) Ug[e] = qnan;
        for (int e = tid; e < n; e += nthr) Wg[e] = qnan;
        if (tid == 0) info[g] = -2;
        return;
    }
    const float* Zg = wk.Zt + (size_t)g * n * n;
    for (int e = tid; e < n * n; e += nthr) {
        const int k = e / n, row = e - k * n;
        zs[k * LDZ + row] = Zg[e];
    }
    const float inv_sc = 1.0f / wk.scale[g];
    for (int k = tid; k < n; k += nthr) lam[k] = wk.d[(size_t)g * n + k] * inv_sc;
    __syncthreads();
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
    // unit columns and the sign rule: the largest-|.| component (lowest row on ties) is positive
    for (int k = w; k < n; k += nwarps) {
        float best = -1.f, best_v = 0.f, ss = 0.f;
        int best_i = 0;
        for (int row = lane; row < n; row += 32) {
            const float v = zs[k * LDZ + row], mag = fabsf(v);
            ss = fmaf(v, v, ss);
            if (mag > best) { best = mag; best_i = row; best_v = v; }
        }
        ss = warp_sum(ss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_v = ov; }
        }
        // ~500 rotations of norm 1 +- 4e-8 each leave the column norms off by a few 1e-6
        const float inv_norm = ss > 0.f ? rsqrtf(ss) : 1.0f;
        if (lane == 0) sgn[k] = (canonical && best_v < 0.f) ? -inv_norm : inv_norm;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nthr) {
        const int row = e / n, c = e - row * n;
        const int k = inv[c];
        Ug[e] = zs[k * LDZ + row] * sgn[k];
    }
    if (tid == 0) info[g] = nit > 0 ? nit : 1;
}

// ==================================================================================== n = 161 .. 288
// The HR-target decomposition (268 nodes, utils/gsr_train.py:33-36) on the same three stages.  A 268-row matrix
// does not fit the register file of one SM (71 824 > 65 536), so these variants keep their matrix in SHARED memory:
//   K1b  the symmetric matrix as its PACKED lower triangle (36 046 floats = 144 KB): thread i owns row i; the
//        mat-vec walks j = 0 .. k-1 reading L[i][j] (j <= i) or L[j][i] (j > i, consecutive over the warp), the
//        rank-2 update touches the i + 1 stored entries of row i; the pivot row k is the tail of the active
//        prefix, contiguous, so nothing is published;
//   K3b  rows of Z are independent: TWO CTAs per graph hold 134 rows each (z[col][row], 145 KB), build their
//        rows of Q from the reflectors and replay the rotation stream with a running register value;
//   K4b  works from the Z^T K3b leaves in global memory (column norms / signs coalesced, 32 x 32 tile transpose).
// K2 is shared.  ~27 ms per 1024 graphs against 136 ms for the Jacobi kernel with its L2-resident workspace.
__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

__global__ void __launch_bounds__(BIGN, 1)
tridiag_packed_kernel(const float* __restrict__ Ain, float* __restrict__ Wout, float* __restrict__ Uout,
                      int32_t* __restrict__ info, const Work wk, int n) {
    extern __shared__ __align__(16) float sm[];
    float* L = sm;                                   // packed lower triangle: L[tri(i) + j], j <= i
    const int ltot = (tri(n) + 3) & ~3;
    float* vbuf = L + ltot;                          // [BIGN]
    float* wbuf = vbuf + BIGN;                       // [BIGN]
    float* pbuf = wbuf + BIGN;                       // [BIGN]
    float* dd = pbuf + BIGN;                         // [BIGN]
    float* ee = dd + BIGN;                           // [BIGN]
    float* red = ee + BIGN;                          // [16]
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nwarps = nthr >> 5;
    const int g = blockIdx.x;
    const float* Ag = Ain + (size_t)g * n * n;

    // ---- load the upper triangle: A_sym[i][j] (j <= i) = A[j][i]; column i of the input, coalesced over i
    int bad = 0;
    float amax = 0.f;
    if (tid < n) {
        for (int j = 0; j <= tid; ++j) {
            const float v = __ldg(Ag + (size_t)j * n + tid);
            if (!isfinite(v)) bad = 1;
            amax = fmaxf(amax, fabsf(v));
            L[tri(tid) + j] = v;
        }
    }
    for (int e = tid; e < 3 * BIGN; e += nthr) vbuf[e] = 0.f;
    if (tid < 16) red[tid] = 0.f;
    if (__syncthreads_or(bad)) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int e = tid; e < n * n; e += nthr) Uout[(size_t)g * n * n + e] = qnan;
        for (int e = tid; e < n; e += nthr) Wout[(size_t)g * n + e] = qnan;
        if (tid == 0) { info[g] = -1; wk.counts[4 * g + 0] = 0; wk.counts[4 * g + 1] = 0; wk.counts[4 * g + 2] = -1; }
        return;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[w] = amax;
    __syncthreads();
    amax = 0.f;
    for (int q = 0; q < nwarps; ++q) amax = fmaxf(amax, red[q]);
    __syncthreads();
    if (lane == 0) red[w] = 0.f;
    float sc = 1.0f;
    if (amax > 0.f && (amax < 9.5367431640625e-07f || amax > 1048576.0f)) sc = ldexpf(1.0f, -max(-120, min(120, ilogbf(amax))));
    if (sc != 1.0f && tid < n) {
        for (int j = 0; j <= tid; ++j) L[tri(tid) + j] *= sc;
    }
    if (tid == 0) { wk.scale[g] = sc; wk.counts[4 * g + 2] = 0; }
    __syncthreads();

    float* Vg = wk.V + (size_t)g * n * LDVB;
    const int i = tid, ti = tri(tid);
    for (int k = n - 1; k >= 1; --k) {
        const float* xr = L + tri(k);                 // row k: columns 0..k (final)
        float ss = 0.f;
        for (int j = lane; j < k - 1; j += 32) { const float x = xr[j]; ss = fmaf(x, x, ss); }
        ss = warp_sum(ss);
        const float alpha = xr[k - 1];
        float tau = 0.f, beta = alpha;
        if (ss != 0.f) {
            const float nrm = sqrtf(fmaf(alpha, alpha, ss));
            beta = -copysignf(nrm, alpha);
            tau = (beta - alpha) / beta;
            const float scl = 1.0f / (alpha - beta);
            const int kpad = (k + 15) & ~15;
            for (int j = lane; j < kpad; j += 32) vbuf[j] = (j < k - 1) ? xr[j] * scl : (j == k - 1 ? 1.0f : 0.0f);   // every warp: same values
        }
        if (tid == 0) { dd[k] = xr[k]; ee[k - 1] = beta; }
        if (tau != 0.f) {                              // uniform
            __syncwarp();
            const bool act = i < k;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            if (act) {
                // A_sym[i][j]: own row up to the diagonal (contiguous), column i of row j beyond it (consecutive over the warp)
                const int kown = min(k, i + 1);
                const float4* V4 = reinterpret_cast<const float4*>(vbuf);
                int j = 0;
                for (; j + 4 <= kown; j += 4) {
                    const float4 v = V4[j >> 2];
                    q0 = fmaf(L[ti + j], v.x, q0); q1 = fmaf(L[ti + j + 1], v.y, q1);
                    q2 = fmaf(L[ti + j + 2], v.z, q2); q3 = fmaf(L[ti + j + 3], v.w, q3);
                }
                for (; j < kown; ++j) q0 = fmaf(L[ti + j], vbuf[j], q0);
                int tj = tri(j) + i;                   // L[j][i], j > i: address advances by j + 1 per row
                for (; j + 4 <= k; j += 4) {
                    const float a0 = L[tj], a1 = L[tj + j + 1], a2 = L[tj + 2 * j + 3], a3 = L[tj + 3 * j + 6];
                    q0 = fmaf(a0, vbuf[j], q0); q1 = fmaf(a1, vbuf[j + 1], q1);
                    q2 = fmaf(a2, vbuf[j + 2], q2); q3 = fmaf(a3, vbuf[j + 3], q3);
                    tj += 4 * j + 10;
                }
                for (; j < k; ++j) { q0 = fmaf(L[tj], vbuf[j], q0); tj += j + 1; }
            }
            const float pi = tau * ((q0 + q1) + (q2 + q3));
            const float vi = act ? vbuf[i] : 0.f;
            const float pvp = warp_sum(act ? pi * vi : 0.f);
            if (lane == 0) red[w] = pvp;
            if (act) pbuf[i] = pi;
            if (w == nwarps - 1) {                     // reflector -> global for K3b (whole 16-column blocks, tau at [288])
                float* vrow = Vg + (size_t)k * LDVB;
                const int kpad = (k + 15) & ~15;
                for (int j = lane; j < kpad; j += 32) vrow[j] = vbuf[j];
                if (lane == 0) vrow[BIGN] = tau;
            }
            __syncthreads();                                                   // (A)
            float pv = 0.f;
            for (int q = 0; q < nwarps; ++q) pv += red[q];
            const float coef = 0.5f * tau * pv;
            const int kpad = (k + 15) & ~15;
            for (int j = lane; j < kpad; j += 32) wbuf[j] = (j < k) ? fmaf(-coef, vbuf[j], pbuf[j]) : 0.0f;        // every warp: same values
            __syncwarp();
            // rank-2 update of the stored triangle, rows dealt to the warps round robin (row a has a + 1 entries:
            // a thread-per-row update would leave the warp of the last rows with twice the average work)
            for (int a = w; a < k; a += nwarps) {
                const float va = vbuf[a], wa = wbuf[a];
                float* La = L + tri(a);
                for (int j = lane; j <= a; j += 32) La[j] = fmaf(-va, wbuf[j], fmaf(-wa, vbuf[j], La[j]));
            }
        } else if (w == nwarps - 1 && lane == 0) {
            Vg[(size_t)k * LDVB + BIGN] = 0.f;
        }
        __syncthreads();                                                       // (B) row k-1 is final
    }
    if (tid == 0) { dd[0] = L[0]; ee[n - 1] = 0.f; Vg[BIGN] = 0.f; }
    __syncthreads();
    for (int j = tid; j < n; j += nthr) {
        wk.d[(size_t)g * n + j] = dd[j];
        wk.e[(size_t)g * n + j] = (j < n - 1) ? ee[j] : 0.f;
    }
}

// K3b: CTA = (graph, row half).  z[col * LDR + r]: column `col` of the CTA's rows (thread = row).
constexpr int RINGB = 3 * 304;                     // floats per warp: 3 reflector rows (292) or 3 x 64 (c, s) pairs

__global__ void __launch_bounds__(160, 1)
tql_apply_smem_kernel(const Work wk, int n, int rows_per_cta) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nwarps = blockDim.x >> 5;
    const int g = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int row0 = half * rows_per_cta, nrows = min(rows_per_cta, n - row0);
    const int LDR = (rows_per_cta + 1) | 1;            // odd pitch with room for the idle lanes' scratch slot
    const int ncols = ((n + 15) & ~15) + 1;            // identity-padded rotations may touch columns n .. 16 ceil(n / 16)
    float* z = sm;
    float* ring = sm + (((size_t)ncols * LDR + 3) & ~(size_t)3) + (size_t)w * RINGB;
    const int* cnt = wk.counts + 4 * g;
    const int nit = cnt[0], nent = cnt[1], status = cnt[2];
    if (status != 0 || nrows <= 0) return;
    const int r = tid < nrows ? tid : nrows;           // idle lanes work on scratch slot `nrows` (< LDR)
    // ---- rows of Q: e_row^T H_{n-1} ... H_1
    for (int c = 0; c < ncols; ++c) z[c * LDR + r] = (tid < nrows && c == row0 + tid) ? 1.0f : 0.0f;
    __syncwarp();
    {
        const float* Vg = wk.V + (size_t)g * n * LDVB;
        auto issue = [&](int k) {
            if (k >= 1) {
                const float* src = Vg + (size_t)k * LDVB;
                float* dst = ring + (k % 3) * 304;
                for (int c = lane; c < LDVB / 4; c += 32) cp_async16(dst + 4 * c, src + 4 * c);
            }
            cp_async_commit();
        };
        issue(n - 1);
        issue(n - 2);
        for (int k = n - 1; k >= 1; --k) {
            cp_async_wait<1>();
            __syncwarp();
            issue(k - 2);
            const float* vr = ring + (k % 3) * 304;
            const float tau = vr[BIGN];
            if (tau != 0.f) {
                const float4* V4 = reinterpret_cast<const float4*>(vr);
                float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
                const int nch = (k + 15) >> 4 << 2;    // whole 16-column blocks: v is zero from column k to the end of its block
#pragma unroll 1
                for (int c = 0; c < nch; c += 4) {     // 16 columns per trip: all 20 loads issued before the first FMA
                    float4 v[4];
                    float zz[16];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = V4[c + u];
                    const float* zp = z + (4 * c) * LDR + r;
#pragma unroll
                    for (int u = 0; u < 16; ++u) zz[u] = zp[u * LDR];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        q0 = fmaf(zz[4 * u], v[u].x, q0); q1 = fmaf(zz[4 * u + 1], v[u].y, q1);
                        q2 = fmaf(zz[4 * u + 2], v[u].z, q2); q3 = fmaf(zz[4 * u + 3], v[u].w, q3);
                    }
                }
                const float s = -tau * ((q0 + q1) + (q2 + q3));
#pragma unroll 1
                for (int c = 0; c < nch; c += 4) {
                    float4 v[4];
                    float zz[16];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = V4[c + u];
                    float* zp = z + (4 * c) * LDR + r;
#pragma unroll
                    for (int u = 0; u < 16; ++u) zz[u] = zp[u * LDR];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        zp[(4 * u) * LDR] = fmaf(s, v[u].x, zz[4 * u]); zp[(4 * u + 1) * LDR] = fmaf(s, v[u].y, zz[4 * u + 1]);
                        zp[(4 * u + 2) * LDR] = fmaf(s, v[u].z, zz[4 * u + 2]); zp[(4 * u + 3) * LDR] = fmaf(s, v[u].w, zz[4 * u + 3]);
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();
    }
    // ---- replay the (identity-padded) rotation stream: iteration = rotations 16 hb - 1 .. 16 lb, descending
    {
        const float2* src = wk.stream + (size_t)g * wk.cap;
        const uint32_t* hdr = wk.hdr + (size_t)g * wk.hdr_cap;
        float2* cring = reinterpret_cast<float2*>(ring);
        const int nblk = (nent + BLK - 1) / BLK;
        auto issue = [&](int b) {
            if (b < nblk) cp_async16(cring + (b % STAGES) * BLK + 2 * lane, src + (size_t)b * BLK + 2 * lane);
            cp_async_commit();
        };
        issue(0);
        issue(1);
        int pos = 0;
        uint32_t h = nit > 0 ? __ldg(hdr) : 0u;
        for (int it = 0; it < nit; ++it) {
            const uint32_t hn = (it + 1 < nit) ? __ldg(hdr + it + 1) : 0u;
            const int lb = (int)(h & 0xffu), hb = (int)((h >> 8) & 0xffu);
            int i = 16 * hb - 1;
            float rv = z[(i + 1) * LDR + r];
            for (int b = hb - 1; b >= lb; --b) {
                if ((pos & (BLK - 1)) == 0) {
                    cp_async_wait<1>();
                    __syncwarp();
                    issue(pos / BLK + 2);
                }
                const float4* cs4 = reinterpret_cast<const float4*>(cring + (pos / BLK % STAGES) * BLK + (pos & (BLK - 1)));
                float* zp = z + i * LDR + r;
                float zi[16];
                float4 q4[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q4[u] = cs4[u];
#pragma unroll
                for (int u = 0; u < 16; ++u) zi[u] = zp[-u * LDR];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    zp[(1 - 2 * u) * LDR] = fmaf(q4[u].x, rv, q4[u].y * zi[2 * u]);
                    rv = fmaf(-q4[u].y, rv, q4[u].x * zi[2 * u]);
                    zp[(-2 * u) * LDR] = fmaf(q4[u].z, rv, q4[u].w * zi[2 * u + 1]);
                    rv = fmaf(-q4[u].w, rv, q4[u].z * zi[2 * u + 1]);
                }
                i -= 16;
                pos += 16;
            }
            z[(i + 1) * LDR + r] = rv;
            h = hn;
        }
        cp_async_wait<0>();
    }
    if (tid < nrows) {
        float* Zg = wk.Zt + (size_t)g * n * n;
        for (int c = 0; c < n; ++c) Zg[(size_t)c * n + row0 + tid] = z[c * LDR + tid];
    }
}

// K4b: like tql_finish_kernel, straight from Z^T in global memory.
__global__ void __launch_bounds__(256)
tql_finish_global_kernel(const Work wk, float* __restrict__ Wout, float* __restrict__ Uout, int32_t* __restrict__ info,
                         int n, int canonical) {
    __shared__ float lam[BIGN], sgn[BIGN];
    __shared__ int inv[BIGN];
    __shared__ float tile[32][33];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nwarps = nthr >> 5;
    const int g = blockIdx.x;
    const int* cnt = wk.counts + 4 * g;
    const int nit = cnt[0], status = cnt[2];
    if (status == -1) return;
    float* Wg = Wout + (size_t)g * n;
    float* Ug = Uout + (size_t)g * n * n;
    if (status != 0) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int e = tid; e < n * n; e += nthr) Ug[e] = qnan;
        for (int e = tid; e < n; e += nthr) Wg[e] = qnan;
        if (tid == 0) info[g] = -2;
        return;
    }
    const float* Zg = wk.Zt + (size_t)g * n * n;
    const float inv_sc = 1.0f / wk.scale[g];
    for (int k = tid; k < n; k += nthr) lam[k] = wk.d[(size_t)g * n + k] * inv_sc;
    __syncthreads();
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
    for (int k = w; k < n; k += nwarps) {
        float best = -1.f, best_v = 0.f, ss = 0.f;
        int best_i = 0;
        for (int row = lane; row < n; row += 32) {
            const float v = Zg[(size_t)k * n + row], mag = fabsf(v);
            ss = fmaf(v, v, ss);
            if (mag > best) { best = mag; best_i = row; best_v = v; }
        }
        ss = warp_sum(ss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_v = ov; }
        }
        const float inv_norm = ss > 0.f ? rsqrtf(ss) : 1.0f;
        if (lane == 0) sgn[k] = (canonical && best_v < 0.f) ? -inv_norm : inv_norm;
    }
    __syncthreads();
    // U[row][c] = Z^T[inv[c]][row] * sgn[inv[c]]: 32 x 32 tiles, read along rows of Z^T, write along rows of U
    const int T = (n + 31) / 32;
    for (int t = 0; t < T * T; ++t) {
        const int rt = t / T, ct = t - rt * T;
        for (int cc = w; cc < 32; cc += nwarps) {
            const int c = ct * 32 + cc, row = rt * 32 + lane;
            float v = 0.f;
            if (c < n && row < n) { const int k = inv[c]; v = Zg[(size_t)k * n + row] * sgn[k]; }
            tile[cc][lane] = v;
        }
        __syncthreads();
        for (int rr = w; rr < 32; rr += nwarps) {
            const int row = rt * 32 + rr, c = ct * 32 + lane;
            if (row < n && c < n) Ug[(size_t)row * n + c] = tile[lane][rr];
        }
        __syncthreads();
    }
    if (tid == 0) info[g] = nit > 0 ? nit : 1;
}

}  // namespace tq
}  // namespace gsr

using namespace gsr;

// K1 with g_eigh_k1_regblocks 16-column blocks of every row in registers
static int launch_k1(const float* A, float* w, float* U, int32_t* info, const gsr::tq::Work& wk, int batch, int threads, int n,
                     int chunk_graphs, cudaStream_t st) {
    using namespace gsr;
    if (g_eigh_k1_regblocks >= 10) tq::tridiag_kernel<10, 2><<<batch, threads, 0, st>>>(A, w, U, info, wk, n, chunk_graphs);
    else if (g_eigh_k1_regblocks == 9) tq::tridiag_kernel<9, 2><<<batch, threads, 0, st>>>(A, w, U, info, wk, n, chunk_graphs);
    else if (g_eigh_k1_regblocks == 8) tq::tridiag_kernel<8, 2><<<batch, threads, 0, st>>>(A, w, U, info, wk, n, chunk_graphs);
    else if (g_eigh_k1_regblocks == 7) tq::tridiag_kernel<7, 3><<<batch, threads, 0, st>>>(A, w, U, info, wk, n, chunk_graphs);     // 3 CTAs per SM: 128 registers
    else tq::tridiag_kernel<6, 3><<<batch, threads, 0, st>>>(A, w, U, info, wk, n, chunk_graphs);
    return launched("eigh_tridiag");
}

// K2 for graphs [g_first, g_end): lane = graph for batches that fill warps, one warp per graph otherwise.  `hog`: ask for all of an
// SM's shared memory so that the CTA runs alone on it (chunked pipeline); wait_count / wait_target: see the kernel.
static int launch_k2(const gsr::tq::Work& wk, int n, int g_first, int g_end, cudaStream_t st, bool hog = false,
                     const int* wait_count = nullptr, int wait_target = 0, int* ctas_launched = nullptr, bool lanes_unchunked = true) {
    using namespace gsr;
    const int count = g_end - g_first;
    // lane = graph: always inside the chunked pipeline (wait_count); for one launch per stage it is never faster than one warp
    // per graph at n <= 160 (tools/eigh_batch_sweep.py), so the segment launcher asks for it only under eigh_k2_lanes >= 2
    if (g_eigh_k2_lanes && (wait_count || (lanes_unchunked && count >= 32))) {
        const size_t per_warp = (size_t)2 * (n + 16) * 32 * sizeof(float);
        const int wpc = (int)std::max<size_t>(1, std::min<size_t>(std::min(tq::K2L_WARPS, std::max(1, g_eigh_k2_warps)), kMaxDynSmem / per_warp));
        const size_t smem = hog ? (size_t)kMaxDynSmem : per_warp * wpc;
        cudaError_t e = cudaFuncSetAttribute(tq::tql_scalar_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
        if (e != cudaSuccess) return fail((int)e, "eigh_tridiag: cudaFuncSetAttribute failed");
        const int ctas = (count + 32 * wpc - 1) / (32 * wpc);
        if (ctas_launched) *ctas_launched += ctas;
        tq::tql_scalar_lanes_kernel<<<ctas, 32 * wpc, smem, st>>>(wk, n, g_first, g_end, g_eigh_ql_max_iter, wait_count, wait_target);
        return launched("eigh_tql_scalar_lanes");
    }
    tq::Work w1 = wk;                                        // the one-warp kernel indexes from 0
    w1.stream += (size_t)g_first * wk.cap; w1.d += (size_t)g_first * n; w1.e += (size_t)g_first * n;
    w1.counts += 4 * (size_t)g_first; w1.hdr += (size_t)g_first * wk.hdr_cap;
    tq::tql_scalar_kernel<<<(count + tq::K2_WARPS - 1) / tq::K2_WARPS, 32 * tq::K2_WARPS, 0, st>>>(w1, n, count, g_eigh_ql_max_iter);
    return launched("eigh_tql_scalar");
}

// Internal streams of the chunked decomposition.  K2 is a latency chain that occupies 1 warp per 32 graphs for ~1.3 ms while K1
// and K3 fill the SMs' register files but leave issue slots and ~10 K registers per SM free: with the batch cut into chunks, each
// on its own stream (K1 -> K2 -> K3 -> K4 per chunk), the chain of one chunk runs under the K1 / K3 of the others.  Earlier chunks
// get the higher stream priority so that the block scheduler drains them first (a pipeline, not a round robin).  The streams fork
// from and join the caller's stream through events, so the whole thing is capturable into a CUDA graph.  Created on first use.
namespace {
constexpr int kMaxChunks = 8;
struct ChunkPool {
    cudaStream_t s[kMaxChunks];
    cudaEvent_t fork, qdone, join[kMaxChunks];
    bool ok = false;
};
ChunkPool* chunk_pool() {
    static ChunkPool pool;
    static bool tried = false;
    if (tried) return pool.ok ? &pool : nullptr;
    tried = true;
    int lo = 0, hi = 0;                                  // lo: least priority (numerically greatest), hi: greatest
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { lo = hi = 0; }
    for (int c = 0; c < kMaxChunks; ++c) {
        const int prio = std::max(hi, std::min(lo, hi + c * std::max(1, (lo - hi + 1)) / kMaxChunks));
        if (cudaStreamCreateWithPriority(&pool.s[c], cudaStreamNonBlocking, prio) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaEventCreateWithFlags(&pool.join[c], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    if (cudaEventCreateWithFlags(&pool.fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&pool.qdone, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    // Lazy module loading: the FIRST launch of a kernel may need a context synchronisation, which never returns while a K2
    // CTA is resident and polling for exactly that kernel (K1) -- so every kernel of the pipeline is loaded here, before any
    // of them can wait for another.
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, gsr::tq::tridiag_kernel<8, 2>) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::tridiag_kernel<9, 2>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, gsr::tq::tridiag_kernel<10, 2>) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::tridiag_kernel<7, 3>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, gsr::tq::tridiag_kernel<6, 3>) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::tql_scalar_lanes_kernel) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, gsr::tq::k2_resident_gate_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::tql_apply_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::form_q_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, gsr::tq::tql_finish_kernel) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    pool.ok = true;
    return &pool;
}
}  // namespace

#ifdef GSR_EIGH_TRACE
// Timeline of the chunked decomposition (tools/eigh_timeline.py builds a private copy of the library with -DGSR_EIGH_TRACE):
// a timing event after every stage of every chunk, read back as milliseconds since the fork.
static cudaEvent_t g_tr_ev[8][5], g_tr_fork;
static bool g_tr_init = false;
static int g_tr_chunk = 0;
static void tr_mark(int stage, cudaStream_t st) { cudaEventRecord(g_tr_ev[g_tr_chunk][stage], st); }
extern "C" __attribute__((visibility("default"))) int gsr_eigh_timeline_read(float* ms, int chunks) {
    for (int c = 0; c < chunks; ++c)
        for (int k = 0; k < 5; ++k)
            if (cudaEventElapsedTime(&ms[c * 5 + k], g_tr_fork, g_tr_ev[c][k]) != cudaSuccess) { cudaGetLastError(); ms[c * 5 + k] = -1.f; }
    return 0;
}
#define GSR_TR(stage) tr_mark(stage, st)
#else
#define GSR_TR(stage) do { } while (0)
#endif

// Kq: Q of graphs [0, batch) into Zt
static int launch_form_q(const gsr::tq::Work& wk, int batch, int threads, int n, cudaStream_t st) {
    using namespace gsr;
    const int upg = threads / 32, units = upg * batch;
    const size_t smemq = (size_t)tq::KQ_SMEM_FLOATS * sizeof(float);
    static const bool attr_ok = cudaFuncSetAttribute(tq::form_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)tq::KQ_SMEM_FLOATS * sizeof(float))) == cudaSuccess;
    if (!attr_ok) { cudaGetLastError(); return fail(GSR_E_UNSUPPORTED, "eigh_tridiag: shared-memory opt-in of the Q kernel failed"); }
    tq::form_q_kernel<<<upg * ((batch + tq::K3_WARPS - 1) / tq::K3_WARPS), 32 * tq::K3_WARPS, smemq, st>>>(wk, n, upg, units);
    return launched("eigh_form_q");
}

// K3 + K4 of the n <= 160 route for graphs [g0, g0 + nb) on stream st
static int launch_small_apply(float* w, float* U, int32_t* info, const gsr::tq::Work& full, int g0, int nb, int n,
                              int canonical_sign, int q_from_zt, cudaStream_t st) {
    using namespace gsr;
    const tq::Work wk = tq::offset_work(full, g0, n);
    U += (size_t)g0 * n * n; w += (size_t)g0 * n; info += g0;
    const int threads = 32 * ((n + 31) / 32);
    const int units_per_graph = threads / 32, total_units = units_per_graph * nb;
    const size_t smem3 = (size_t)tq::K3_WARPS * tq::RING_FLOATS * sizeof(float);
    GSR_TR(2);
    tq::tql_apply_kernel<<<(total_units + tq::K3_WARPS - 1) / tq::K3_WARPS, 32 * tq::K3_WARPS, smem3, st>>>(wk, n, units_per_graph, total_units, q_from_zt);
    int rc = launched("eigh_tql_apply");
    if (rc) return rc;
    GSR_TR(3);
    const size_t smem4 = ((size_t)n * tq::LDZ + 3 * (size_t)n) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tq::tql_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
    if (e != cudaSuccess) return fail((int)e, "eigh_tridiag: cudaFuncSetAttribute failed");
    tq::tql_finish_kernel<<<nb, 256, smem4, st>>>(wk, w, U, info, n, canonical_sign);
    rc = launched("eigh_tql_finish");
    GSR_TR(4);
    return rc;
}

// One segment (<= 1024 graphs) of the n <= 160 route; A, w, U, info and wk point at the segment's first graph.
static int eigh_small_segment(const float* A, float* w, float* U, int32_t* info, const gsr::tq::Work& wk, int batch, int n,
                              int canonical_sign, cudaStream_t st) {
    using namespace gsr;
    const int threads = 32 * ((n + 31) / 32);
    int rc = 0;
    // Chunked pipeline (384 graphs or more; measured, `tools/eigh_batch_sweep.py`: one launch per stage with the one-warp QL kernel is faster below:
    // 2.25 against 2.62 ms for 256 graphs, 2.69 against 2.66 for 384, 5.09 against 3.92 for 1024): K1 is ONE launch on the caller's stream (its CTAs run in graph order and count the
    // finished graphs of each chunk); every chunk has a stream with K2 -> K3 -> K4.  The K2 CTAs are launched FIRST: each asks
    // for a whole SM, becomes resident at the fork and polls its chunk's count, so the ~1.3 ms rotation chain of a chunk runs
    // on its own SM (one warp per scheduler, nothing else to share issue slots with) under the K1 / K3 of the other chunks,
    // which have the remaining SMs.  K1 never waits for anything, so the polling cannot deadlock; if a K2 CTA is placed late
    // (no empty SM at the fork) it simply starts late.  Forks from / joins the caller's stream through events: capturable.
    const int per = std::max(128, (((batch + kMaxChunks - 1) / kMaxChunks) + 127) & ~127);       // graphs per chunk
    int nchunks = (g_eigh_chunks > 1 && g_eigh_k2_lanes && batch >= 384) ? (batch + per - 1) / per : 1;
    // the resident K2 CTAs need K1 to run NEXT to them: not when every launch blocks the host until its kernel has finished.
    // (Profilers that serialise kernels -- ncu, compute-sanitizer -- need gsr_set_tuning("eigh_chunks", 1) for the same reason;
    // a K2 CTA that never sees its count traps after ~4 s instead of hanging the GPU.)
    static const bool launch_blocking = [] { const char* v = getenv("CUDA_LAUNCH_BLOCKING"); return v && atoi(v) != 0; }();
    if (launch_blocking) nchunks = 1;
    ChunkPool* pool = nchunks > 1 ? chunk_pool() : nullptr;
    if (!pool || nchunks > std::min(g_eigh_chunks, kMaxChunks)) nchunks = 1;
#ifdef GSR_EIGH_TRACE
    if (!g_tr_init) {
        for (int c = 0; c < 8; ++c) for (int k = 0; k < 5; ++k) cudaEventCreate(&g_tr_ev[c][k]);
        cudaEventCreate(&g_tr_fork);
        g_tr_init = true;
    }
    cudaEventRecord(g_tr_fork, st);
    g_tr_chunk = 0;
#endif
    if (nchunks == 1) {
        GSR_TR(0);
        if ((rc = launch_k1(A, w, U, info, wk, batch, threads, n, 0, st))) return rc;
        GSR_TR(1);
        if ((rc = launch_k2(wk, n, 0, batch, st, false, nullptr, 0, nullptr, g_eigh_k2_lanes >= 2))) return rc;      // one warp per graph unless forced
        // Q by its own launch here too (eigh_q_launch = 0: K3 forms it, in row form): Kq's column form does 2/3 of the flops,
        // and every batch size forms Q with the same arithmetic (a shard of a batch reproduces the batch's eigenvectors)
        const int qown = g_eigh_q_launch >= 1 ? 1 : 0;
        if (qown && (rc = launch_form_q(wk, batch, threads, n, st))) return rc;
        if ((rc = launch_small_apply(w, U, info, wk, 0, batch, n, canonical_sign, qown, st))) return rc;
    } else {
        if (cudaMemsetAsync(wk.chunk_done, 0, 64 * sizeof(int), st) != cudaSuccess) return fail((int)cudaGetLastError(), "eigh_tridiag: cudaMemsetAsync failed");
        if (cudaEventRecord(pool->fork, st) != cudaSuccess) return fail((int)cudaGetLastError(), "eigh_tridiag: cudaEventRecord failed");
        int forked = 0, k2_ctas = 0;
        const int qmode = g_eigh_q_launch ? 1 : 0;
        for (int c = 0; c < nchunks && !rc; ++c) {
            const int g0 = c * per, g1 = std::min(batch, g0 + per);
            if (cudaStreamWaitEvent(pool->s[c], pool->fork, 0) != cudaSuccess) { rc = fail((int)cudaGetLastError(), "eigh_tridiag: cudaStreamWaitEvent failed"); break; }
            forked = c + 1;
#ifdef GSR_EIGH_TRACE
            g_tr_chunk = c; tr_mark(1, pool->s[c]);
#endif
            rc = launch_k2(wk, n, g0, g1, pool->s[c], true, wk.chunk_done + c, g1 - g0, &k2_ctas);
        }
        if (!rc) {
            tq::k2_resident_gate_kernel<<<1, 1, 0, st>>>(wk.chunk_done + 32, k2_ctas);
            rc = launched("eigh_k2_gate");
        }
        if (!rc) {
#ifdef GSR_EIGH_TRACE
            g_tr_chunk = 0;
#endif
            GSR_TR(0);
            rc = launch_k1(A, w, U, info, wk, batch, threads, n, per, st);
        }
        if (!rc && qmode) {
            // Q of all graphs right behind K1 (it runs while the first chunks' rotation chains are still going); the chunk
            // streams wait for it before their K3
            rc = launch_form_q(wk, batch, threads, n, st);
            if (!rc && cudaEventRecord(pool->qdone, st) != cudaSuccess) rc = fail((int)cudaGetLastError(), "eigh_tridiag: cudaEventRecord failed");
        }
        for (int c = 0; c < nchunks && !rc; ++c) {
            const int g0 = c * per, g1 = std::min(batch, g0 + per);
#ifdef GSR_EIGH_TRACE
            g_tr_chunk = c;
#endif
            if (qmode && cudaStreamWaitEvent(pool->s[c], pool->qdone, 0) != cudaSuccess) { rc = fail((int)cudaGetLastError(), "eigh_tridiag: cudaStreamWaitEvent failed"); break; }
            rc = launch_small_apply(w, U, info, wk, g0, g1 - g0, n, canonical_sign, qmode ? 1 : 0, pool->s[c]);
        }
        // (on an error the streams forked so far are still joined, so that a capture is not left forked)
        for (int c = 0; c < forked; ++c) { cudaEventRecord(pool->join[c], pool->s[c]); cudaStreamWaitEvent(st, pool->join[c], 0); }
        if (rc) return rc;
    }
    return 0;
}

extern "C" size_t gsr_eigh_tridiag_workspace(int batch, int n) {
    if (batch <= 0 || n <= 0 || n > tq::BIGN) return 0;
    return tq::carve(nullptr, nullptr, batch, n);
}

extern "C" int gsr_eigh_tridiag_f32(const float* A, float* w, float* U, int32_t* info, void* workspace,
                                    size_t workspace_bytes, int batch, int n, int canonical_sign, void* stream) {
    if (!A || !w || !U || !info || batch < 0 || n <= 0) return fail(GSR_E_ARG, "eigh_tridiag: bad argument");
    if (batch == 0) return 0;
    if (n > tq::BIGN) return fail(GSR_E_UNSUPPORTED, "eigh_tridiag: n > 288 is not supported");
    if (!workspace || workspace_bytes < tq::carve(nullptr, nullptr, batch, n))
        return fail(GSR_E_ARG, "eigh_tridiag: workspace missing or smaller than gsr_eigh_tridiag_workspace(batch, n)");
    if (!aligned16(workspace)) return fail(GSR_E_ALIGN, "eigh_tridiag: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    tq::Work wk;
    tq::carve(&wk, workspace, batch, n);
    const int threads = 32 * ((n + 31) / 32);

    if (n > tq::MAXN) {
        // ---- 161 .. 288 nodes: shared-memory variants
        const size_t smem1 = ((((size_t)n * (n + 1) / 2 + 3) & ~(size_t)3) + 5 * tq::BIGN + 16) * sizeof(float);
        cudaError_t e1 = cudaFuncSetAttribute(tq::tridiag_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e1 != cudaSuccess) return fail((int)e1, "eigh_tridiag: cudaFuncSetAttribute failed");
        tq::tridiag_packed_kernel<<<batch, threads, smem1, st>>>(A, w, U, info, wk, n);
        int rcb = launched("eigh_tridiag_packed");
        if (rcb) return rcb;
        if ((rcb = launch_k2(wk, n, 0, batch, st))) return rcb;
        const int rows_per_cta = (n + 1) / 2, thr3 = 32 * ((rows_per_cta + 31) / 32);
        const int ldr = (rows_per_cta + 1) | 1, ncols = ((n + 15) & ~15) + 1;
        const size_t smem3b = ((((size_t)ncols * ldr + 3) & ~(size_t)3) + (size_t)(thr3 / 32) * tq::RINGB) * sizeof(float);
        e1 = cudaFuncSetAttribute(tq::tql_apply_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3b);
        if (e1 != cudaSuccess) return fail((int)e1, "eigh_tridiag: cudaFuncSetAttribute failed");
        tq::tql_apply_smem_kernel<<<2 * batch, thr3, smem3b, st>>>(wk, n, rows_per_cta);
        if ((rcb = launched("eigh_tql_apply_smem"))) return rcb;
        tq::tql_finish_global_kernel<<<batch, 256, 0, st>>>(wk, w, U, info, n, canonical_sign);
        if ((rcb = launched("eigh_tql_finish_global"))) return rcb;
        // QL failures (info == -2) are redone on the device by the Jacobi kernel; its scratch for a matrix that does
        // not fit shared memory is the graph's own (now dead) rotation stream: 2 cap >= 4 n^2 floats
        return g_eigh_redo ? eigh_jacobi_redo_failed(A, w, U, info, reinterpret_cast<float*>(wk.stream), 2LL * wk.cap, batch, n, canonical_sign, st) : 0;
    }

    // segments of at most 1024 graphs, one after the other (a pipeline over more graphs would give more than 16 SMs to the QL
    // chains: 2048 graphs in one pipeline 8.7 ms, in two 7.8)
    const int nseg = (batch + 1023) / 1024, seg = (((batch + nseg - 1) / nseg) + 127) & ~127;
    for (int g0 = 0; g0 < batch; g0 += seg) {
        const int nb = std::min(seg, batch - g0);
        const int rc = eigh_small_segment(A + (size_t)g0 * n * n, w + (size_t)g0 * n, U + (size_t)g0 * n * n, info + g0,
                                          tq::offset_work(wk, g0, n), nb, n, canonical_sign, st);
        if (rc) return rc;
    }
    return g_eigh_redo ? eigh_jacobi_redo_failed(A, w, U, info, reinterpret_cast<float*>(wk.stream), 2LL * wk.cap, batch, n, canonical_sign, st) : 0;
}
