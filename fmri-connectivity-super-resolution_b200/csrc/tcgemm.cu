// Batched GEMM family on the 5th-generation tensor cores (tcgen05, sm_100a) at fp32 accuracy.
//
// The contractions of the GSR-Net path (models/layers.py:37,45,54,59,85-86 and their autograd)
// must hold fp32 tolerances end to end, so a single TF32 pass (10-bit mantissa) is not enough.
// Every fp32 operand x is split by the producer warps into hi = tf32(x) and lo = tf32(x - hi)
// (both round-to-nearest) and the product is accumulated as  lo.hi + hi.lo + hi.hi  by three
// `tcgen05.mma.kind::tf32` passes into one fp32 accumulator in tensor memory (the dropped lo.lo
// term is 2^-22 relative).  Per CTA:
//   * tile 128 (M) x up to 288 (N = 160 + up to 128, two UMMA N ranges) x 32 (K per stage), so
//     one 268-column subject row block needs a single N tile;
//   * 8 producer warps read the fp32 operands from global memory (float4, any layout: the
//     transposes of the backward pass are absorbed by the UMMA K-major / MN-major descriptor
//     bit), split them and store hi / lo tiles in UMMA canonical layouts: K-major tiles use
//     the no-swizzle layout (8 x 16 B core matrices), MN-major tiles the 128B_BASE32B layout
//     (the only one the hardware takes for MN-major 32-bit operands: 4 k x 128 B atoms whose
//     32-byte blocks are XOR-permuted by the k-row index); a warp writes 512 contiguous bytes
//     per instruction, conflict free;
//   * 1 MMA warp: an elected lane issues 4 k-steps x 3 passes x 1-2 N ranges per stage, then
//     `tcgen05.commit` releases the stage (mbarrier) back to the producers; 2 stages;
//   * epilogue: the 8 producer warps read the accumulator with `tcgen05.ld.32x32b.x8`
//     (warp w: TMEM lanes 32*(w%4).., column half w/4), apply alpha / bias / addend and store.
// Weight gradients (sum over subjects) keep accumulating in TMEM across the subjects dealt to
// the CTA; partial tiles are then reduced in a fixed order (same scheme as the SIMT kernel).
#include "common.cuh"

namespace gsr {
namespace tc {

constexpr int BM = 128;
constexpr int BN = 288;                 // 160 + up to 128 (multiple of the 32-element MN atom)
constexpr int BN0 = 160;
constexpr int BK = 32;
constexpr int STAGES = 2;
constexpr int PRODUCER_WARPS = 8;
constexpr int THREADS = 32 * (PRODUCER_WARPS + 1);
constexpr int A_TILE = BM * BK * 4;     // bytes of one tf32 A tile (hi or lo)
constexpr int B_TILE = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;            // 102400
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 64;           // + barriers, tmem pointer
constexpr int TMEM_COLS = 512;
constexpr int A_PATCHES = BM * BK / 128;                        // 32 patches of 128 floats (one per warp instruction)
constexpr int B_PATCHES = BN * BK / 128;                        // 72
constexpr int PATCHES = A_PATCHES + B_PATCHES;                  // 104
constexpr int PATCH_ITERS = (PATCHES + PRODUCER_WARPS - 1) / PRODUCER_WARPS;   // 13

struct Params {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    long long sA, sB, sC;
    int batch, splits;
    float alpha;
    int epilogue;
    const float* extra; int ld_extra; long long s_extra;
    int mn_swap;
};

int g_mn_swap = 0;      // bring-up knob (unused)

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fault (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (sm_100 version bit set).  layout_type: 0 = SWIZZLE_NONE
// ("interleave"), 1 = SWIZZLE_128B_BASE32B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}
// Instruction descriptor: fp32 accumulate, tf32 x tf32, M = 128, runtime N, operand majors.
__device__ __forceinline__ uint32_t instr_desc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---------------------------------------------------------------- operand staging
// One operand tile is ROWS (M or N extent) x BK.
// KCONTIG (stored [row][k] in global memory, k contiguous) -> UMMA K-major, no swizzle:
//   core matrix (row/8, k/4) = 8 rows x 16 B at (row/8) * SBO + (k/4) * LBO, LBO = 128,
//   SBO = (BK/4) * 128 = 1024.  A warp instruction covers 8 rows x 16 k (4 core matrices).
// else (stored [k][row], row contiguous) -> UMMA MN-major, SWIZZLE_128B_BASE32B:
//   atom (row/32, k/4) = 4 k-rows x 128 B (32 consecutive rows) at (row/32) * LBO + (k/4) * SBO,
//   LBO = 512, SBO = (ROWS/32) * 512; inside a k-row the four 32-byte blocks are XOR-permuted
//   by the k-row index: block (row%32)/8 of k-row k%4 sits at block index ((row%32)/8) ^ (k%4)
//   (byte-address bits [5,7) ^= bits [7,9): measured on B200 with the probe of tools/tc_probe.py,
//   the layout is not documented in this environment).  A warp instruction covers one atom.
template <int ROWS, bool KCONTIG>
__device__ __forceinline__ void patch_coords(int patch, int lane, int& row, int& k, uint32_t& soff) {
    if (KCONTIG) {
        const int i8 = lane & 7, c4 = lane >> 3;              // row of the core matrix, 16-byte chunk of the patch
        const int r8 = patch >> 1, k16 = patch & 1;           // patch = 8 rows x 16 k
        row = r8 * 8 + i8;
        k = k16 * 16 + c4 * 4;
        soff = (uint32_t)(r8 * ((BK / 4) * 128) + (k >> 2) * 128 + i8 * 16);
    } else {
        const int c8 = lane & 7, k4i = lane >> 3;             // 16-byte chunk of the k-row, k-row of the atom
        const int r32 = patch % (ROWS / 32), k4 = patch / (ROWS / 32);
        k = k4 * 4 + k4i;
        row = r32 * 32 + c8 * 4;
        soff = (uint32_t)(r32 * 512 + k4 * ((ROWS / 32) * 512) + k4i * 128 + (((c8 >> 1) ^ k4i) << 5) + ((c8 & 1) << 4));
    }
}

template <int ROWS, bool KCONTIG>
__device__ __forceinline__ float4 load_patch(const float* __restrict__ base, int ld, int rows_valid, int K,
                                             int row0, int k0, int patch, int lane, uint32_t& soff) {
    int row, k;
    patch_coords<ROWS, KCONTIG>(patch, lane, row, k, soff);
    row += row0; k += k0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
        if (row < rows_valid && k < K) {
            const float* p = base + (size_t)row * ld + k;
            if (k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
            else { v.x = __ldg(p); if (k + 1 < K) v.y = __ldg(p + 1); if (k + 2 < K) v.z = __ldg(p + 2); }
        }
    } else {
        if (k < K && row < rows_valid) {
            const float* p = base + (size_t)k * ld + row;
            if (row + 3 < rows_valid) v = __ldg(reinterpret_cast<const float4*>(p));
            else { v.x = __ldg(p); if (row + 1 < rows_valid) v.y = __ldg(p + 1); if (row + 2 < rows_valid) v.z = __ldg(p + 2); }
        }
    }
    return v;
}

template <bool SWIZZLE32>
__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t soff, float4 v) {
    float4 h, l;
    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(hi_tile + soff) = h;
    *reinterpret_cast<float4*>(lo_tile + soff) = l;
}

// ---------------------------------------------------------------- the kernel
// TA: op(A) = A^T (A stored [K][M], MN-major);  TB: op(B) = B^T (B stored [N][K], K-major).
template <bool TA, bool TB, bool REDUCE>
__global__ void __launch_bounds__(THREADS, 1)
tcgemm_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // full[2], empty[2], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES), bar_accum = smem_u32(bars + 2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ncols = min(BN, ((p.N - n0) + 15) & ~15);          // UMMA N extent of this tile (multiple of 16)
    const int n_first = ncols > 256 ? BN0 : ncols;               // first UMMA N range
    const int n_second = ncols - n_first;                        // second one (0 or 128)
    const int nkb = (p.K + BK - 1) / BK;
    const int b_first = blockIdx.z;
    const int b_step = REDUCE ? p.splits : p.batch;
    const int nsub = REDUCE ? (p.batch - b_first + b_step - 1) / b_step : 1;
    const int total_kb = nsub * nkb;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, PRODUCER_WARPS * 32); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accum, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < PRODUCER_WARPS) {
        // ===== producers: global fp32 -> (hi, lo) tf32 tiles in shared memory
        for (int it = 0; it < total_kb; ++it) {
            const int s = it % STAGES, use = it / STAGES;
            const int b = b_first + (it / nkb) * b_step, kb = it % nkb;
            const float* A = p.A + (size_t)b * p.sA;
            const float* B = p.B + (size_t)b * p.sB;
            float4 v[PATCH_ITERS];
            uint32_t so[PATCH_ITERS];
#pragma unroll
            for (int i = 0; i < PATCH_ITERS; ++i) {
                const int patch = warp + i * PRODUCER_WARPS;
                if (patch < A_PATCHES) v[i] = load_patch<BM, !TA>(A, p.lda, p.M, p.K, m0, kb * BK, patch, lane, so[i]);
                else if (patch < PATCHES) v[i] = load_patch<BN, TB>(B, p.ldb, p.N, p.K, n0, kb * BK, patch - A_PATCHES, lane, so[i]);
            }
            if (use > 0) mbar_wait(bar_empty + 8 * s, (use - 1) & 1);      // MMA has drained this stage
            uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < PATCH_ITERS; ++i) {
                const int patch = warp + i * PRODUCER_WARPS;
                if (p.mn_swap) {
                    // bring-up probe: A (K-major) selects k = m % 8, every word of the B tile holds
                    // (part of) its own word index, so C[m][n] names the word read as B(n, m % 8)
                    if (patch < A_PATCHES) {
                        int row, k; uint32_t so2;
                        patch_coords<BM, true>(patch, lane, row, k, so2);
                        float4 a;
                        a.x = (k + 0 == (row & 7)) ? 1.f : 0.f; a.y = (k + 1 == (row & 7)) ? 1.f : 0.f;
                        a.z = (k + 2 == (row & 7)) ? 1.f : 0.f; a.w = (k + 3 == (row & 7)) ? 1.f : 0.f;
                        *reinterpret_cast<float4*>(st + so2) = a;
                        *reinterpret_cast<float4*>(st + A_TILE + so2) = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else if (patch < PATCHES) {
                        const uint32_t w = so[i] >> 2;
                        float4 b;
                        if (p.mn_swap == 1) b = make_float4((float)((w + 0) & 1023u), (float)((w + 1) & 1023u), (float)((w + 2) & 1023u), (float)((w + 3) & 1023u));
                        else b = make_float4((float)((w + 0) >> 10), (float)((w + 1) >> 10), (float)((w + 2) >> 10), (float)((w + 3) >> 10));
                        *reinterpret_cast<float4*>(st + 2 * A_TILE + so[i]) = b;
                        *reinterpret_cast<float4*>(st + 2 * A_TILE + B_TILE + so[i]) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    continue;
                }
                if (patch < A_PATCHES) store_split<TA>(st, st + A_TILE, so[i], v[i]);
                else if (patch < PATCHES) store_split<!TB>(st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, so[i], v[i]);
            }
            fence_async_smem();                                            // generic-proxy writes -> async proxy (UMMA)
            mbar_arrive(bar_full + 8 * s);
        }
    } else {
        // ===== MMA issuer: one elected lane
        if (lane == 0) {
            const uint32_t idesc0 = instr_desc(n_first, TA ? 1 : 0, TB ? 0 : 1);
            const uint32_t idesc1 = instr_desc(n_second > 0 ? n_second : 16, TA ? 1 : 0, TB ? 0 : 1);
            // strides of the canonical layouts (bytes), the k-step (K = 8) advance of the start
            // address and the start of the second N range (BN0 rows further) inside a B tile
            const uint32_t a_lt = TA ? 1u : 0u, b_lt = TB ? 0u : 1u;
            const uint32_t a_lbo = TA ? 512 : 128, a_sbo = TA ? (BM / 32) * 512 : (BK / 4) * 128;
            const uint32_t b_lbo = TB ? 128 : 512, b_sbo = TB ? (BK / 4) * 128 : (BN / 32) * 512;
            const uint32_t a_kstep = TA ? 2 * (BM / 32) * 512 : 256, b_kstep = TB ? 256 : 2 * (BN / 32) * 512;
            const uint32_t b_n1 = TB ? (BN0 / 8) * ((BK / 4) * 128) : (BN0 / 32) * 512;
            for (int it = 0; it < total_kb; ++it) {
                const int s = it % STAGES, use = it / STAGES;
                mbar_wait(bar_full + 8 * s, use & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t a_hi = st, a_lo = st + A_TILE, b_hi = st + 2 * A_TILE, b_lo = st + 2 * A_TILE + B_TILE;
#pragma unroll
                for (int j = 0; j < BK / 8; ++j) {
                    const uint64_t dah = smem_desc(a_hi + j * a_kstep, a_lbo, a_sbo, a_lt);
                    const uint64_t dal = smem_desc(a_lo + j * a_kstep, a_lbo, a_sbo, a_lt);
                    const uint64_t dbh = smem_desc(b_hi + j * b_kstep, b_lbo, b_sbo, b_lt);
                    const uint64_t dbl = smem_desc(b_lo + j * b_kstep, b_lbo, b_sbo, b_lt);
                    const uint32_t first = (it == 0 && j == 0) ? 0u : 1u;
                    umma_tf32(tmem_base, dal, dbh, idesc0, first);          // small terms first
                    umma_tf32(tmem_base, dah, dbl, idesc0, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc0, 1u);
                    if (n_second > 0) {
                        const uint64_t dbh1 = smem_desc(b_hi + b_n1 + j * b_kstep, b_lbo, b_sbo, b_lt);
                        const uint64_t dbl1 = smem_desc(b_lo + b_n1 + j * b_kstep, b_lbo, b_sbo, b_lt);
                        umma_tf32(tmem_base + BN0, dal, dbh1, idesc1, first);
                        umma_tf32(tmem_base + BN0, dah, dbl1, idesc1, 1u);
                        umma_tf32(tmem_base + BN0, dah, dbh1, idesc1, 1u);
                    }
                }
                umma_commit(bar_empty + 8 * s);                              // stage free once these MMAs retire
            }
            umma_commit(bar_accum);                                          // accumulator complete
        }
        __syncwarp();
    }

    // ===== epilogue: TMEM -> registers -> global (producer warps)
    if (warp < PRODUCER_WARPS) {
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const int row = m0 + q * 32 + lane;
        const int c_begin = half * (BN / 2), c_end = min(c_begin + BN / 2, ncols);
        float* C;
        int ldc;
        if (REDUCE) { C = p.C + (size_t)blockIdx.z * p.M * p.N; ldc = p.N; }
        else        { C = p.C + (size_t)blockIdx.z * p.sC;      ldc = p.ldc; }
        const float* extra = (!REDUCE && p.epilogue == GSR_EPI_ADD) ? p.extra + (size_t)blockIdx.z * p.s_extra : nullptr;
        const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
        for (int c = c_begin; c < c_end; c += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);     // whole warp, convergent
            if (row < p.M) {
                const int n = n0 + c;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (!REDUCE) {
                        v[i] *= p.alpha;
                        if (n + i < p.N) {
                            if (p.epilogue == GSR_EPI_BIAS) v[i] += __ldg(p.extra + n + i);
                            else if (p.epilogue == GSR_EPI_ADD) v[i] += __ldg(extra + (size_t)row * p.ld_extra + n + i);
                        }
                    }
                }
                float* dst = C + (size_t)row * ldc + n;
                if (vec && n + 7 < p.N) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (n + i < p.N) dst[i] = v[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCER_WARPS) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <bool REDUCE>
static int launch(const Params& p, int transA, int transB, dim3 grid, cudaStream_t st) {
#define GSR_TC_LAUNCH(TA, TB)                                                                                     \
    do {                                                                                                          \
        auto k = tcgemm_kernel<TA, TB, REDUCE>;                                                                   \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);         \
        if (e != cudaSuccess) return fail((int)e, "tcgemm: cudaFuncSetAttribute failed");                        \
        k<<<grid, THREADS, SMEM_BYTES, st>>>(p);                                                                  \
    } while (0)
    if (!transA && !transB) GSR_TC_LAUNCH(false, false);
    else if (!transA && transB) GSR_TC_LAUNCH(false, true);
    else if (transA && !transB) GSR_TC_LAUNCH(true, false);
    else GSR_TC_LAUNCH(true, true);
#undef GSR_TC_LAUNCH
    return launched(REDUCE ? "tcgemm_reduce" : "tcgemm");
}

static bool operands_ok(const Params& p) {
    return aligned16(p.A) && aligned16(p.B) && (p.lda % 4 == 0) && (p.ldb % 4 == 0) && (p.sA % 4 == 0) && (p.sB % 4 == 0);
}

// C[m][n] = alpha * sum_z ws[z][m][n] (+ extra[m][n]); fixed summation order.
__global__ void reduce_partials_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                       int splits, float alpha, const float* __restrict__ extra, int ld_extra) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * N) return;
    const int m = e / N, n = e - m * N;
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += __ldg(ws + (size_t)z * M * N + e);
    acc *= alpha;
    if (extra) acc += __ldg(extra + (size_t)m * ld_extra + n);
    C[(size_t)m * ldc + n] = acc;
}

}  // namespace tc
}  // namespace gsr

using namespace gsr;

extern "C" __attribute__((visibility("default"))) int gsr_tcgemm_debug(int key, int value) {
    if (key == 0) tc::g_mn_swap = value;
    return 0;
}

extern "C" int gsr_tcgemm_supported(int M, int N, int K, int lda, int ldb) {
    return (M > 0 && N > 0 && K > 0 && lda % 4 == 0 && ldb % 4 == 0) ? 1 : 0;
}

extern "C" int gsr_tcgemm_f32(const float* A, const float* B, float* C,
                              int M, int N, int K, int lda, int ldb, int ldc,
                              long long strideA, long long strideB, long long strideC,
                              int transA, int transB, int batch, float alpha,
                              int epilogue, const float* extra, int ld_extra, long long stride_extra,
                              void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K < 0 || batch < 0) return fail(GSR_E_ARG, "tcgemm: bad argument");
    if (epilogue != GSR_EPI_NONE && !extra) return fail(GSR_E_ARG, "tcgemm: epilogue needs `extra`");
    if (epilogue < 0 || epilogue > GSR_EPI_ADD) return fail(GSR_E_ARG, "tcgemm: unknown epilogue");
    if (M == 0 || N == 0 || batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "tcgemm: batch > 65535");
    tc::Params p{A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, 1, alpha,
                 epilogue, extra, ld_extra, stride_extra, tc::g_mn_swap};
    if (!tc::operands_ok(p)) return fail(GSR_E_ALIGN, "tcgemm: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    dim3 grid((M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN, batch);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "tcgemm: N too large");
    return tc::launch<false>(p, transA, transB, grid, (cudaStream_t)stream);
}

extern "C" size_t gsr_tcgemm_reduce_workspace(int M, int N, int batch, int* splits_out) {
    const int tiles = ((N + tc::BN - 1) / tc::BN) * ((M + tc::BM - 1) / tc::BM);
    int splits = (kNumSMs + tiles - 1) / tiles;          // one CTA per SM (TMEM is allocated whole)
    if (splits > batch) splits = batch;
    if (splits < 1) splits = 1;
    if (splits_out) *splits_out = splits;
    return (size_t)splits * (size_t)M * (size_t)N;
}

extern "C" int gsr_tcgemm_reduce_f32(const float* A, const float* B, float* C,
                                     int M, int N, int K, int lda, int ldb, int ldc,
                                     long long strideA, long long strideB,
                                     int transA, int transB, int batch, float alpha,
                                     int epilogue, const float* extra, int ld_extra,
                                     float* workspace, int splits, void* stream) {
    if (!A || !B || !C || !workspace || M <= 0 || N <= 0 || K < 0 || batch <= 0 || splits <= 0)
        return fail(GSR_E_ARG, "tcgemm_reduce: bad argument");
    if (epilogue != GSR_EPI_NONE && epilogue != GSR_EPI_ADD) return fail(GSR_E_ARG, "tcgemm_reduce: epilogue must be NONE or ADD");
    if (epilogue == GSR_EPI_ADD && !extra) return fail(GSR_E_ARG, "tcgemm_reduce: epilogue needs `extra`");
    if (splits > batch) splits = batch;
    tc::Params p{A, B, workspace, M, N, K, lda, ldb, N, strideA, strideB, 0, batch, splits, 1.0f,
                 GSR_EPI_NONE, nullptr, 0, 0, tc::g_mn_swap};
    if (!tc::operands_ok(p)) return fail(GSR_E_ALIGN, "tcgemm_reduce: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    dim3 grid((M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN, splits);
    int rc = tc::launch<true>(p, transA, transB, grid, (cudaStream_t)stream);
    if (rc) return rc;
    const int total = M * N;
    tc::reduce_partials_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        workspace, C, M, N, ldc, splits, alpha, epilogue == GSR_EPI_ADD ? extra : nullptr, ld_extra);
    return launched("tcgemm_reduce_partials");
}
