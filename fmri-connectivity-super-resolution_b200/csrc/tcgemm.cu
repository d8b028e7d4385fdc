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
//   * 14 producer warps (two groups alternating k-blocks) read the fp32 operands from global memory (float4, any layout: the
//     transposes of the backward pass are absorbed by the UMMA K-major / MN-major descriptor
//     bit), split them and store hi / lo tiles in UMMA canonical layouts: K-major tiles use
//     the no-swizzle layout (8 x 16 B core matrices), MN-major tiles the 128B_BASE32B layout
//     (the only one the hardware takes for MN-major 32-bit operands: 4 k x 128 B atoms whose
//     32-byte blocks are XOR-permuted by the k-row index); a warp writes 512 contiguous bytes
//     per instruction, conflict free;
//   * 1 MMA warp: an elected lane issues 4 k-steps x 3 passes x 1-2 N ranges per stage, then
//     `tcgen05.commit` releases the stage (mbarrier) back to the producers; 2 stages;
//   * epilogue: 12 warps read the accumulator with `tcgen05.ld.32x32b.x32` (warp w: TMEM lanes
//     32*(w%4).., column group w/4) into a row-major staging tile in the idle operand buffers;
//     all warps then apply alpha / bias / addend and write full coalesced rows.
// Weight gradients (sum over subjects) keep accumulating in TMEM across the subjects dealt to
// the CTA; partial tiles are then reduced in a fixed order (same scheme as the SIMT kernel).
#include "tcgemm.cuh"

namespace gsr {
namespace tc {

int g_pairs = 1;        // gsr_set_tuning("tcgemm_pairs"): CTA-pair kernel (cta_group::2) for gsr_tcgemm_f32 (default), 0 = one CTA per tile
int g_dual = 1;         // gsr_set_tuning("tcgemm_dual"): gsr_tcgemm_dual_f32 accumulates both products in tensor memory (default), 0 = two launches, the second with the addend epilogue
int g_tail = 1;         // gsr_set_tuning("tcgemm_ragged_rows"): FP32 kernels for the <= 16 ragged rows after the CTA pairs (default), 0 = one-CTA tensor kernel

#ifdef GSR_TC_TRACE
// Pipeline trace (tools/tc_trace.py builds a private copy of the library with -DGSR_TC_TRACE):
// one CTA stamps %globaltimer at its pipeline events; slot layout [role][k-block][event].
__device__ unsigned long long g_trace[4 * 16 * 8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_STAMP(role, kb, ev) do { if (blockIdx.x == 1 && blockIdx.z == 7 && (kb) < 16) g_trace[((role) * 16 + (kb)) * 8 + (ev)] = gtime(); } while (0)
#else
#define TC_STAMP(role, kb, ev) do { } while (0)
#endif

constexpr int BN = 288;                 // 160 + up to 128 (multiple of the 32-element MN atom)
constexpr int BN0 = 160;
constexpr int STAGES = 2;
constexpr int PRODUCER_WARPS = 14;    // two groups of 7 (+ 1 MMA warp = 480 threads -> 128 registers each)
constexpr int GROUP_WARPS = PRODUCER_WARPS / 2;
constexpr int EPILOGUE_WARPS = 12;    // 3 column groups x 4 TMEM lane quarters
constexpr int THREADS = 32 * (PRODUCER_WARPS + 1);
constexpr int A_TILE = BM * BK * 4;     // bytes of one tf32 A tile (hi or lo)
constexpr int B_TILE = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;            // 102400
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 64;           // + barriers, tmem pointer
constexpr int TMEM_COLS = 512;
constexpr int A_PATCHES = BM * BK / 128;                        // 32 patches of 128 floats (one per warp instruction)
constexpr int B_PATCHES = BN * BK / 128;                        // 72
constexpr int PATCHES = A_PATCHES + B_PATCHES;                  // 104
constexpr int PATCH_ITERS = (PATCHES + GROUP_WARPS - 1) / GROUP_WARPS;   // 15: one group stages a whole k-block

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
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, GROUP_WARPS * 32); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accum, 1);
        fence_barrier_init();
    }
    __syncthreads();                                                      // barriers initialised
    // The tensor-memory allocation (MMA warp only, below) overlaps with the producers' address set-up and their
    // first loads: nobody but the MMA thread needs the accumulator address before the accumulator barrier fires.
    uint32_t tmem_base = 0;

    if (warp < PRODUCER_WARPS) {
        // ===== producers: global fp32 -> (hi, lo) tf32 tiles in shared memory.
        // The producer warps form two groups; group g stages the k-blocks it % 2 == g into stage g.
        // A group's iteration is load -> split -> store -> proxy fence -> arrive, and the fence
        // (MEMBAR) waits for every outstanding access of its thread, so a thread cannot keep the
        // next k-block's loads in flight across it; instead the OTHER group's loads are in
        // flight while this one splits and stores.  Per-thread patch state (global offset,
        // validity, shared offset) is computed once; the k loop only advances pointers.
        static_assert(STAGES == 2, "one stage per producer group");
        const int grp = warp / GROUP_WARPS, gw = warp % GROUP_WARPS;
        long long goff[PATCH_ITERS];        // element offset of this thread's float4 at k-block 0
        uint32_t so[PATCH_ITERS];           // byte offset inside the (hi or lo) tile
        uint32_t ok_mask = 0;               // bit i: the float4 of patch i lies inside the operand (rows)
        uint32_t dead_mask = 0;             // bit i: an A patch whose rows all lie below the operand (ragged last M tile,
                                            //        12 of 128 rows for the 268-row subject matrices): neither loaded nor
                                            //        staged -- accumulator rows are independent and those rows are never stored
        const int rows_valid = p.M - m0;
#pragma unroll
        for (int i = 0; i < PATCH_ITERS; ++i) {
            const int patch = gw + i * GROUP_WARPS;
            int row = 0, k = 0;
            goff[i] = 0; so[i] = 0;
            if (patch < A_PATCHES) {
                patch_coords<BM, !TA>(patch, lane, row, k, so[i]);
                if ((TA ? (patch % (BM / 32)) * 32 : (patch >> 1) * 8) >= rows_valid) dead_mask |= 1u << i;
                row += m0;
                goff[i] = TA ? (long long)k * p.lda + row : (long long)row * p.lda + k;
                if (TA ? (row + 3 < p.M) : (row < p.M)) ok_mask |= 1u << i;
            } else if (patch < PATCHES) {
                patch_coords<BN, TB>(patch - A_PATCHES, lane, row, k, so[i]);
                row += n0;
                goff[i] = TB ? (long long)row * p.ldb + k : (long long)k * p.ldb + row;
                if (TB ? (row < p.N) : (row + 3 < p.N)) ok_mask |= 1u << i;
            }
        }
        const long long a_kadv = TA ? (long long)BK * p.lda : BK, b_kadv = TB ? BK : (long long)BK * p.ldb;
        uint8_t* st = smem + grp * STAGE_BYTES;
        for (int it = grp; it < total_kb; it += 2) {
            const int use = it / STAGES;
            const int b = b_first + (it / nkb) * b_step, kb = it % nkb;
            const float* A = p.A + (size_t)b * p.sA;
            const float* B = p.B + (size_t)b * p.sB;
            float4 v[PATCH_ITERS];
            if (gw == 0 && lane == 0) TC_STAMP(grp, it, 0);                 // loads issued from here
            if ((kb + 1) * BK <= p.K) {                                    // whole k-block inside K: no k predicates
#pragma unroll
                for (int i = 0; i < PATCH_ITERS; ++i) {
                    const int patch = gw + i * GROUP_WARPS;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (patch < PATCHES && ((ok_mask >> i) & 1u)) {
                        const float* src = patch < A_PATCHES ? A + goff[i] + kb * a_kadv : B + goff[i] + kb * b_kadv;
                        v[i] = __ldg(reinterpret_cast<const float4*>(src));
                    }
                }
            } else {                                                       // ragged tail (k and row edges): generic path
                uint32_t dummy;
#pragma unroll
                for (int i = 0; i < PATCH_ITERS; ++i) {
                    const int patch = gw + i * GROUP_WARPS;
                    if (patch < A_PATCHES) { if (!((dead_mask >> i) & 1u)) v[i] = load_patch<BM, !TA>(A, p.lda, p.M, p.K, m0, kb * BK, patch, lane, dummy); }
                    else if (patch < PATCHES) v[i] = load_patch<BN, TB>(B, p.ldb, p.N, p.K, n0, kb * BK, patch - A_PATCHES, lane, dummy);
                }
            }
            if (gw == 0 && lane == 0) TC_STAMP(grp, it, 1);                 // loads in flight, waiting for the stage
            if (use > 0) mbar_wait(bar_empty + 8 * grp, (use - 1) & 1);    // MMA has drained this stage
            if (gw == 0 && lane == 0) TC_STAMP(grp, it, 2);                 // stage free
#ifdef GSR_TC_TRACE
            if (gw == 0 && lane == 0) { volatile float sink = v[PATCH_ITERS - 1].x + v[0].x; (void)sink; TC_STAMP(grp, it, 3); }   // loads landed
#endif
#pragma unroll
            for (int i = 0; i < PATCH_ITERS; ++i) {
                const int patch = gw + i * GROUP_WARPS;
                if (patch < A_PATCHES) { if (!((dead_mask >> i) & 1u)) store_split(st, st + A_TILE, so[i], v[i]); }
                else if (patch < PATCHES) store_split(st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, so[i], v[i]);
            }
            if (gw == 0 && lane == 0) TC_STAMP(grp, it, 4);                 // split + stores issued
            fence_async_smem();                                            // generic-proxy writes -> async proxy (UMMA)
            if (gw == 0 && lane == 0) TC_STAMP(grp, it, 5);                 // fence done
            mbar_arrive(bar_full + 8 * grp);
        }
    } else {
        // ===== MMA issuer: one elected lane
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
        tc_fence_before();
        __syncwarp();
        tc_fence_after();
        tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
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
            const uint32_t a_dhi = desc_hi(a_sbo, a_lt), b_dhi = desc_hi(b_sbo, b_lt);
            const uint32_t a_step = a_kstep >> 4, b_step = b_kstep >> 4;     // descriptor address units (16 B)
            uint32_t base_lo[STAGES][4];                                      // A hi, A lo, B hi, B lo of every stage
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                base_lo[s][0] = desc_lo(st, a_lbo);
                base_lo[s][1] = desc_lo(st + A_TILE, a_lbo);
                base_lo[s][2] = desc_lo(st + 2 * A_TILE, b_lbo);
                base_lo[s][3] = desc_lo(st + 2 * A_TILE + B_TILE, b_lbo);
            }
            const uint32_t n1 = b_n1 >> 4;
            uint32_t acc = 0;                                                 // 0 only for the very first k-step
            for (int it = 0; it < total_kb; ++it) {
                const int s = it % STAGES, use = it / STAGES;
                TC_STAMP(2, it, 0);                                          // MMA thread waits for the stage
                mbar_wait(bar_full + 8 * s, use & 1);
                TC_STAMP(2, it, 1);                                          // stage full
                tc_fence_after();
                uint32_t ah = base_lo[s][0], al = base_lo[s][1], bh = base_lo[s][2], bl = base_lo[s][3];
#pragma unroll
                for (int j = 0; j < BK / 8; ++j) {
                    umma_tf32(tmem_base, al, a_dhi, bh, b_dhi, idesc0, acc);          // small terms first
                    umma_tf32(tmem_base, ah, a_dhi, bl, b_dhi, idesc0, 1u);
                    umma_tf32(tmem_base, ah, a_dhi, bh, b_dhi, idesc0, 1u);
                    if (n_second > 0) {
                        umma_tf32(tmem_base + BN0, al, a_dhi, bh + n1, b_dhi, idesc1, acc);
                        umma_tf32(tmem_base + BN0, ah, a_dhi, bl + n1, b_dhi, idesc1, 1u);
                        umma_tf32(tmem_base + BN0, ah, a_dhi, bh + n1, b_dhi, idesc1, 1u);
                    }
                    acc = 1u;
                    ah += a_step; al += a_step; bh += b_step; bl += b_step;
                }
                umma_commit(bar_empty + 8 * s);                              // stage free once these MMAs retire
                TC_STAMP(2, it, 2);                                          // MMAs issued + committed
            }
            umma_commit(bar_accum);                                          // accumulator complete
        }
        __syncwarp();
    }

    // ===== epilogue: TMEM -> registers -> shared staging tile -> global.
    // tcgen05.ld hands every thread one accumulator ROW (lane = row), so storing straight to
    // global memory would scatter 16-byte pieces over 32 rows per instruction.  The tile is
    // staged row-major in the (now idle) operand buffers with an odd pitch (conflict-free both
    // ways) and written out by all warps as full, coalesced rows.
    constexpr int CPITCH = BN + 4;                                          // 292 floats: float4 aligned, 4 (mod 32) -> a quarter
                                                                            // warp of row-owning lanes hits 32 distinct banks
    float* Cs = reinterpret_cast<float*>(smem);
    if (warp < EPILOGUE_WARPS) {
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
        const int q = warp & 3, part = warp >> 2;                          // TMEM lane quarter, column group
        constexpr int COLS_PER_PART = BN / (EPILOGUE_WARPS / 4);            // 96
        const int r = q * 32 + lane;
        const int c_begin = part * COLS_PER_PART;
        const int c_end = (q * 32 < p.M - m0) ? min(c_begin + COLS_PER_PART, ncols) : c_begin;   // no output rows in this quarter
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);   // whole warp, convergent
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<uint4*>(Cs + r * CPITCH + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        float* C;
        int ldc;
        if (REDUCE) { C = p.C + (size_t)blockIdx.z * p.sC; ldc = p.N; }
        else        { C = p.C + (size_t)blockIdx.z * p.sC;      ldc = p.ldc; }
        const float* extra = (!REDUCE && p.epilogue == GSR_EPI_ADD) ? p.extra + (size_t)blockIdx.z * p.s_extra : nullptr;
        const int rows = min(BM, p.M - m0), cols = min(BN, p.N - n0);
        const int epi = REDUCE ? GSR_EPI_NONE : p.epilogue;
        const float alpha = REDUCE ? 1.0f : p.alpha;
        const bool vec = ((cols & 3) == 0) && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) &&
                         (epi == GSR_EPI_NONE || ((reinterpret_cast<uintptr_t>(p.extra) & 15u) == 0 &&
                                                  (epi != GSR_EPI_ADD || ((p.ld_extra & 3) == 0 && (p.s_extra & 3) == 0))));
        if (vec) {
            // float4 per thread over the flattened (row, 4-column group) space: full 16-byte stores, rows stay contiguous
            const int c4n = cols >> 2;
            for (int e = tid; e < rows * c4n; e += THREADS) {
                const int r = e / c4n, c = (e - r * c4n) << 2;
                float4 v = *reinterpret_cast<const float4*>(Cs + r * CPITCH + c);
                v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
                if (epi == GSR_EPI_BIAS) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.extra + n0 + c));
                    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                } else if (epi == GSR_EPI_ADD) {
                    const float4 a4 = *reinterpret_cast<const float4*>(extra + (size_t)(m0 + r) * p.ld_extra + n0 + c);     // plain load: the addend may alias C
                    v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
                }
                *reinterpret_cast<float4*>(C + (size_t)(m0 + r) * ldc + n0 + c) = v;
            }
        } else {
        for (int r = warp; r < rows; r += PRODUCER_WARPS + 1) {
            const float* src = Cs + r * CPITCH;
            float* dst = C + (size_t)(m0 + r) * ldc + n0;
            for (int c = lane; c < cols; c += 32) {
                float v = src[c] * alpha;
                if (epi == GSR_EPI_BIAS) v += __ldg(p.extra + n0 + c);
                else if (epi == GSR_EPI_ADD) v += extra[(size_t)(m0 + r) * p.ld_extra + n0 + c];
                dst[c] = v;
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCER_WARPS) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------- the CTA-pair kernel (tcgen05 cta_group::2)
// Two CTAs on the two SMs of a TPC compute a 256-row block: each stages its own 128 rows of A but only HALF of
// the B operand, and the leader issues M = 256 MMAs that read both halves (tools/probes/tc2_probe.cu established
// on the hardware that thread-staged canonical tiles work and that CTA rank r supplies B rows r N/2 ..).  B is
// 69 % of the staging work of a 128 x 288 tile, so the pair moves 35 % less through the load / split / store path
// that bounds tcgemm_kernel (profiles/r1_tcgemm_pipeline_trace.txt).
//   * K-major B (op(B) = B^T): N ranges 160 + 112, the pair member holds rows [80 r, 80 r + 80) and
//     [160 + 56 r, 160 + 56 r + 56) -- 136 local rows;
//   * MN-major B: the 128B_BASE32B atoms are 32 columns wide, so the ranges are 192 + 128 (320 >= 268, the
//     padding costs MMA time only) and the member holds columns [96 r, 96 r + 96) and [192 + 64 r, .. + 64).
//   * full[s] lives in the leader and counts the producers of BOTH CTAs (the peer arrives through DSMEM with
//     release.cluster); tcgen05.commit multicasts the stage release and the accumulator-ready signal to both.
constexpr int B2_ROWS_K = 136, B2_ROWS_MN = 160;
constexpr int B2_TILE_MAX = B2_ROWS_MN * BK * 4;                 // 20480
constexpr int STAGE2_BYTES = 2 * A_TILE + 2 * B2_TILE_MAX;       // 73728
constexpr int STAGES2 = 2;                                       // three stages (216 KB) measured slower: they leave no L1 for the operand loads
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 128;
constexpr int CPITCH2 = 276;                                     // epilogue staging pitch (128 x 276 floats = 141 KB of the 216 KB of stages)

template <bool TA, bool TB, bool REDUCE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tcgemm2_kernel(const Params p) {
    constexpr int BROWS = TB ? B2_ROWS_K : B2_ROWS_MN;            // local B tile rows (N extent held by this CTA)
    constexpr int R0 = TB ? 80 : 96, R1 = TB ? 56 : 64;           // per-CTA share of the two N ranges
    constexpr int NR0 = 2 * R0, NR1 = 2 * R1;                     // UMMA N of the two ranges: 160 + 112 / 192 + 128
    constexpr int B2_TILE = BROWS * BK * 4;
    constexpr int B2_PATCHES = BROWS * BK / 128;                  // 34 / 40
    constexpr int PATCHES2 = A_PATCHES + B2_PATCHES;              // 66 / 72
    constexpr int ITERS2 = (PATCHES2 + GROUP_WARPS - 1) / GROUP_WARPS;   // 10 / 11
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);  // full[3] (leader's are used), empty[3], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES2), bar_accum = smem_u32(bars + 2 * STAGES2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int m0 = (blockIdx.x) * BM;                             // blockIdx.x = 2 pair + rank: consecutive M tiles
    const int nkb = (p.K + BK - 1) / BK;
    // REDUCE (weight gradients): the pair accumulates the subjects blockIdx.z, blockIdx.z + splits, ... in TMEM
    const int b_first = blockIdx.z;
    const int b_step = REDUCE ? p.splits : p.batch;
    const int nsub = REDUCE ? (p.batch - b_first + b_step - 1) / b_step : 1;
    const int total_kb = nsub * nkb;

    if (tid == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(bar_full + 8 * s, 2 * GROUP_WARPS * 32); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accum, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCER_WARPS) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < PRODUCER_WARPS) {
        const int grp = warp / GROUP_WARPS, gw = warp % GROUP_WARPS;
        int goff[ITERS2];
        uint32_t so[ITERS2];
        uint32_t ok_mask = 0, dead_mask = 0;
        const int rows_valid = p.M - m0;
#pragma unroll
        for (int i = 0; i < ITERS2; ++i) {
            const int patch = gw + i * GROUP_WARPS;
            int row = 0, k = 0;
            goff[i] = 0; so[i] = 0;
            if (patch < A_PATCHES) {
                patch_coords<BM, !TA>(patch, lane, row, k, so[i]);
                if ((TA ? (patch % (BM / 32)) * 32 : (patch >> 1) * 8) >= rows_valid) dead_mask |= 1u << i;
                row += m0;
                goff[i] = TA ? k * p.lda + row : row * p.lda + k;
                if (TA ? (row + 3 < p.M) : (row < p.M)) ok_mask |= 1u << i;
            } else if (patch < PATCHES2) {
                patch_coords<BROWS, TB>(patch - A_PATCHES, lane, row, k, so[i]);
                // local row / column of the half tile -> column n of op(B)
                const int n = row < R0 ? R0 * (int)rank + row : NR0 + R1 * (int)rank + (row - R0);
                goff[i] = TB ? n * p.ldb + k : k * p.ldb + n;
                if (TB ? (n < p.N) : (n + 3 < p.N)) ok_mask |= 1u << i;
            }
        }
        const int a_kadv = TA ? BK * p.lda : BK, b_kadv = TB ? BK : BK * p.ldb;
        const uint32_t full_leader0 = map_to_cta(bar_full, 0);
        int sub = 0, kb = grp;                                         // k-block `it` = (subject sub of this pair, k-block kb)
        for (int it = grp; it < total_kb; it += 2) {                   // group g stages the k-blocks g, g + 2, ...
            while (kb >= nkb) { kb -= nkb; ++sub; }
            const float* A = p.A + (size_t)(b_first + sub * b_step) * p.sA;
            const float* B = p.B + (size_t)(b_first + sub * b_step) * p.sB;
            const int sidx = it % STAGES2, use = it / STAGES2;
            uint8_t* st = smem + sidx * STAGE2_BYTES;
            float4 v[ITERS2];
            const bool whole = (kb + 1) * BK <= p.K;
#pragma unroll
            for (int i = 0; i < ITERS2; ++i) {
                const int patch = gw + i * GROUP_WARPS;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (patch < PATCHES2 && ((ok_mask >> i) & 1u)) {
                    const float* src = patch < A_PATCHES ? A + goff[i] + kb * a_kadv : B + goff[i] + kb * b_kadv;
                    const bool kmajor = patch < A_PATCHES ? !TA : TB;          // float4 runs along k
                    if (whole) v[i] = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        // ragged last k-block: K % 4 == 0 here (checked by the launcher), so a float4 along k is all in or all
                        // out; along the rows the whole k-row is in or out
                        int row, k; uint32_t dummy;
                        if (patch < A_PATCHES) patch_coords<BM, !TA>(patch, lane, row, k, dummy);
                        else patch_coords<BROWS, TB>(patch - A_PATCHES, lane, row, k, dummy);
                        (void)kmajor;
                        if (kb * BK + k < p.K) v[i] = __ldg(reinterpret_cast<const float4*>(src));
                    }
                }
            }
            if (use > 0) mbar_wait(bar_empty + 8 * sidx, (use - 1) & 1);
#pragma unroll
            for (int i = 0; i < ITERS2; ++i) {
                const int patch = gw + i * GROUP_WARPS;
                if (patch < A_PATCHES) { if (!((dead_mask >> i) & 1u)) store_split(st, st + A_TILE, so[i], v[i]); }
                else if (patch < PATCHES2) store_split(st + 2 * A_TILE, st + 2 * A_TILE + B2_TILE, so[i], v[i]);
            }
            fence_async_smem();
            mbar_arrive_cluster(full_leader0 + 8 * sidx);          // the leader's full[stage] counts both CTAs' producers
            kb += 2;
        }
    } else {
        if (rank == 0 && lane == 0) {
            const uint32_t idesc0 = instr_desc2(NR0, TA ? 1 : 0, TB ? 0 : 1);
            const uint32_t idesc1 = instr_desc2(NR1, TA ? 1 : 0, TB ? 0 : 1);
            const uint32_t a_lt = TA ? 1u : 0u, b_lt = TB ? 0u : 1u;
            const uint32_t a_lbo = TA ? 512 : 128, a_sbo = TA ? (BM / 32) * 512 : (BK / 4) * 128;
            const uint32_t b_lbo = TB ? 128 : 512, b_sbo = TB ? (BK / 4) * 128 : (BROWS / 32) * 512;
            const uint32_t a_kstep = TA ? 2 * (BM / 32) * 512 : 256, b_kstep = TB ? 256 : 2 * (BROWS / 32) * 512;
            const uint32_t b_n1 = TB ? (R0 / 8) * ((BK / 4) * 128) : (R0 / 32) * 512;
            const uint32_t a_dhi = desc_hi(a_sbo, a_lt), b_dhi = desc_hi(b_sbo, b_lt);
            const uint32_t a_step = a_kstep >> 4, b_step = b_kstep >> 4, n1 = b_n1 >> 4;
            uint32_t acc = 0;
            for (int it = 0; it < total_kb; ++it) {
                const int s = it % STAGES2, use = it / STAGES2;
                mbar_wait_cluster(bar_full + 8 * s, use & 1);
                tc_fence_after();
                const uint32_t stb = smem_u32(smem + s * STAGE2_BYTES);
                uint32_t ah = desc_lo(stb, a_lbo), al = desc_lo(stb + A_TILE, a_lbo);
                uint32_t bh = desc_lo(stb + 2 * A_TILE, b_lbo), bl = desc_lo(stb + 2 * A_TILE + B2_TILE, b_lbo);
#pragma unroll
                for (int j = 0; j < BK / 8; ++j) {
                    umma2_tf32(tmem_base, al, a_dhi, bh, b_dhi, idesc0, acc);
                    umma2_tf32(tmem_base, ah, a_dhi, bl, b_dhi, idesc0, 1u);
                    umma2_tf32(tmem_base, ah, a_dhi, bh, b_dhi, idesc0, 1u);
                    umma2_tf32(tmem_base + NR0, al, a_dhi, bh + n1, b_dhi, idesc1, acc);
                    umma2_tf32(tmem_base + NR0, ah, a_dhi, bl + n1, b_dhi, idesc1, 1u);
                    umma2_tf32(tmem_base + NR0, ah, a_dhi, bh + n1, b_dhi, idesc1, 1u);
                    acc = 1u;
                    ah += a_step; al += a_step; bh += b_step; bl += b_step;
                }
                umma_commit2(bar_empty + 8 * s);
            }
            umma_commit2(bar_accum);
        }
        __syncwarp();
    }

    // ===== epilogue (each CTA drains its own 128 accumulator rows; accumulator column c = column n of the product)
    float* Cs = reinterpret_cast<float*>(smem);
    const int ncols = min(p.N, NR0 + NR1);
    if (warp < EPILOGUE_WARPS) {
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const int q = warp & 3, part = warp >> 2;
        constexpr int COLS_PER_PART = 96;
        const int r = q * 32 + lane;
        const int c_begin = part * COLS_PER_PART;
        const int c_end = (q * 32 < p.M - m0) ? min(c_begin + COLS_PER_PART, (ncols + 31) & ~31) : c_begin;
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                if (c0 + i < CPITCH2 - 4) *reinterpret_cast<uint4*>(Cs + r * CPITCH2 + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        float* C = p.C + (size_t)blockIdx.z * p.sC;                 // REDUCE: sC = M * N, the pitch of the partials
        const int ldc = REDUCE ? p.N : p.ldc;
        const float* extra = (!REDUCE && p.epilogue == GSR_EPI_ADD) ? p.extra + (size_t)blockIdx.z * p.s_extra : nullptr;
        const int rows = min(BM, p.M - m0), cols = ncols;
        const int epi = REDUCE ? GSR_EPI_NONE : p.epilogue;
        const float alpha = REDUCE ? 1.0f : p.alpha;
        const int c4n = cols >> 2;                                   // the launcher guarantees N % 4 == 0 and aligned operands
        for (int e = tid; e < rows * c4n; e += THREADS) {
            const int r = e / c4n, c = (e - r * c4n) << 2;
            float4 v = *reinterpret_cast<const float4*>(Cs + r * CPITCH2 + c);
            v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
            if (epi == GSR_EPI_BIAS) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.extra + c));
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            } else if (epi == GSR_EPI_ADD) {
                const float4 a4 = *reinterpret_cast<const float4*>(extra + (size_t)(m0 + r) * p.ld_extra + c);
                v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
            }
            *reinterpret_cast<float4*>(C + (size_t)(m0 + r) * ldc + c) = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // the peer may still be reading this CTA's B half / TMEM pair
    if (warp == PRODUCER_WARPS) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// Pairs cover the first 2 * floor(tiles_m / 2) M tiles; a last odd tile (the 12 ragged rows of a 268-row subject
// matrix) goes through tcgemm_kernel on an offset view.
static bool pair_ok(const Params& p) {
    return p.N <= 268 + 4 && p.N % 4 == 0 && p.K % 4 == 0 && p.M >= 2 * BM && p.ldc % 4 == 0 && aligned16(p.C) && p.sC % 4 == 0 &&
           (p.epilogue == GSR_EPI_NONE || (aligned16(p.extra) && (p.epilogue != GSR_EPI_ADD || (p.ld_extra % 4 == 0 && p.s_extra % 4 == 0))));
}

template <bool REDUCE>
static int launch_pairs(const Params& p, int transA, int transB, int zdim, cudaStream_t st) {
    const int tiles_m = (p.M + BM - 1) / BM, pairs = tiles_m / 2;
    dim3 grid(2 * pairs, 1, zdim);
#define GSR_TC2_LAUNCH(TA, TB)                                                                                   \
    do {                                                                                                          \
        auto k = tcgemm2_kernel<TA, TB, REDUCE>;                                                                  \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);        \
        if (e != cudaSuccess) return fail((int)e, "tcgemm2: cudaFuncSetAttribute failed");                       \
        k<<<grid, THREADS, SMEM2_BYTES, st>>>(p);                                                                 \
    } while (0)
    if (!transA && !transB) GSR_TC2_LAUNCH(false, false);
    else if (!transA && transB) GSR_TC2_LAUNCH(false, true);
    else if (transA && !transB) GSR_TC2_LAUNCH(true, false);
    else GSR_TC2_LAUNCH(true, true);
#undef GSR_TC2_LAUNCH
    return launched(REDUCE ? "tcgemm2_reduce" : "tcgemm2");
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

// ---------------------------------------------------------------- the ragged last rows of M (M % 128 <= 16), FP32 FMAs
// The 12 rows that 268 leaves after two 128-row tiles cost the tensor kernel 70 % of a full tile (all of B is
// split and staged for 4.5 % of the rows: 150-160 us per 1024 subjects).  For row-major B they are a memory-shaped
// product instead: the CTA keeps the <= 16 rows of op(A) in shared memory (k-major, 16 floats per k, quad-swizzled
// so that the one-off transposing store is at most 4-way conflicted), streams B through a cp.async double buffer
// (a warp per k row, 16-byte copies), and each thread owns two columns n, n + 136 with all rows in registers: per k
// two scalar LDS, R4 broadcast LDS.128 and 8 R4 FFMA.  Plain FP32 accumulation in k order (the 3xTF32 products of
// the tensor kernels differ from it by rounding only; tests compare both against float64).  80 us per 1024
// subjects.  op(B) = B^T has its own kernel below (a transposing variant of this one measured 215 us); the
// subject-reduced products stay on the tensor kernel (a reducing variant of this kernel with 4 x 148 splits was
// measured twice, before and after the staging was tuned: 0.363 against 0.349 ms for the whole product).
namespace tail {
constexpr int T_THREADS = 128, T_COLS = 272, T_KC = 16, T_KMAX = 288, T_ROWS = 16;      // columns: tid, tid + 128; warp 0 also 256 + (tid & 15)
constexpr int T_SMEM_FLOATS = T_KMAX * T_ROWS + 2 * T_KC * T_COLS;

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool TA, int R4>
__global__ void __launch_bounds__(T_THREADS, 4) tail_kernel(const Params p) {
    extern __shared__ __align__(16) float tsm[];
    float* As = tsm;                                   // [k][16], quad q of row k at ((q ^ ((k >> 1) & 3)) << 2)
    float* Bs = tsm + T_KMAX * T_ROWS;                 // [2][T_KC][T_COLS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = p.M, K = p.K, N = p.N;
    // subjects in REVERSE order: the tensor kernel that has just run on the same operands walked them upwards, so the
    // last ones are the part of B still resident in the 126 MB L2
    const int bz = (int)gridDim.z - 1 - (int)blockIdx.z;
    const int nchunks = (K + T_KC - 1) / T_KC, kpad = nchunks * T_KC;
    const float* A = p.A + (size_t)bz * p.sA;
    const float* B = p.B + (size_t)bz * p.sB;

    auto copy_b = [&](int c) {                         // a warp per k row; rows past K and columns past N are zero filled
        float* dst = Bs + (c & 1) * T_KC * T_COLS;
        const int k0 = c * T_KC;
        for (int kk = warp; kk < T_KC; kk += T_THREADS / 32) {
            const bool row_ok = k0 + kk < K;
            const float* src = B + (size_t)(row_ok ? k0 + kk : 0) * p.ldb;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n4 = (lane + 32 * j) * 4;
                if (n4 < T_COLS) { const bool ok = row_ok && n4 < N; cp_async16(dst + kk * T_COLS + n4, src + (ok ? n4 : 0), ok); }
            }
        }
        cp_commit();
    };
    copy_b(0);
    // op(A) rows, zero filled to 4 R4 rows and to whole chunks of k (all loads of a thread are issued before its
    // first store: one memory round trip, not one per row)
    if (!TA) {
        constexpr int KJ = (T_KMAX + T_THREADS - 1) / T_THREADS;
        float v[R4 * 4][KJ];
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r)
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                const int k = tid + j * T_THREADS;
                v[r][j] = (r < rows && k < K) ? __ldg(A + (size_t)r * p.lda + k) : 0.f;
            }
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r)
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                const int k = tid + j * T_THREADS;
                if (k < kpad) As[k * T_ROWS + ((((r >> 2) ^ (k >> 1)) & 3) << 2) + (r & 3)] = v[r][j];
            }
    } else {
        constexpr int PASSES = (T_KMAX + T_THREADS / 16 - 1) / (T_THREADS / 16);       // 29 k rows per thread
        const int r = tid & 15;
        float v[PASSES];
#pragma unroll
        for (int j = 0; j < PASSES; ++j) {
            const int k = (tid >> 4) + j * (T_THREADS / 16);
            v[j] = (r < rows && k < K) ? __ldg(A + (size_t)k * p.lda + r) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < PASSES; ++j) {
            const int k = (tid >> 4) + j * (T_THREADS / 16);
            if (r < R4 * 4 && k < kpad) As[k * T_ROWS + ((((r >> 2) ^ (k >> 1)) & 3) << 2) + (r & 3)] = v[j];
        }
    }
    cp_wait<0>();
    __syncthreads();

    // Four full warps own columns tid and tid + 128; the 16 columns from 256 on (the 12 that 268 leaves) are a third column of
    // warp 0, in a loop of its own behind a warp-uniform branch (the first version gave them to a fifth warp with 8 live lanes
    // that issued the whole loop: a fifth of the kernel's instructions for 4 % of its outputs).
    float acc[R4 * 4][2], acc2[R4 * 4];
#pragma unroll
    for (int r = 0; r < R4 * 4; ++r) acc[r][0] = acc[r][1] = acc2[r] = 0.f;
    const bool third = warp == 0 && N > 2 * T_THREADS;
    for (int c = 0; c < nchunks; ++c) {
        // chunk c + 1 goes to the buffer whose last readers (chunk c - 1) all passed the barrier of iteration c - 1
        if (c + 1 < nchunks) copy_b(c + 1);
        {
            const float* bs = Bs + (c & 1) * T_KC * T_COLS + tid;
            const float* as = As + c * T_KC * T_ROWS;
#pragma unroll
            for (int kk = 0; kk < T_KC; ++kk) {
                const float b0 = bs[kk * T_COLS], b1 = bs[kk * T_COLS + T_THREADS];
#pragma unroll
                for (int q = 0; q < R4; ++q) {
                    const float4 a = *reinterpret_cast<const float4*>(as + kk * T_ROWS + (((q ^ (kk >> 1)) & 3) << 2));
                    acc[4 * q + 0][0] = fmaf(a.x, b0, acc[4 * q + 0][0]); acc[4 * q + 0][1] = fmaf(a.x, b1, acc[4 * q + 0][1]);
                    acc[4 * q + 1][0] = fmaf(a.y, b0, acc[4 * q + 1][0]); acc[4 * q + 1][1] = fmaf(a.y, b1, acc[4 * q + 1][1]);
                    acc[4 * q + 2][0] = fmaf(a.z, b0, acc[4 * q + 2][0]); acc[4 * q + 2][1] = fmaf(a.z, b1, acc[4 * q + 2][1]);
                    acc[4 * q + 3][0] = fmaf(a.w, b0, acc[4 * q + 3][0]); acc[4 * q + 3][1] = fmaf(a.w, b1, acc[4 * q + 3][1]);
                }
            }
            if (third) {                                                   // warp uniform
                const float* bt = Bs + (c & 1) * T_KC * T_COLS + 2 * T_THREADS + (tid & 15);
#pragma unroll 4
                for (int kk = 0; kk < T_KC; ++kk) {
                    const float b2 = bt[kk * T_COLS];
#pragma unroll
                    for (int q = 0; q < R4; ++q) {
                        const float4 a = *reinterpret_cast<const float4*>(as + kk * T_ROWS + (((q ^ (kk >> 1)) & 3) << 2));
                        acc2[4 * q + 0] = fmaf(a.x, b2, acc2[4 * q + 0]); acc2[4 * q + 1] = fmaf(a.y, b2, acc2[4 * q + 1]);
                        acc2[4 * q + 2] = fmaf(a.z, b2, acc2[4 * q + 2]); acc2[4 * q + 3] = fmaf(a.w, b2, acc2[4 * q + 3]);
                    }
                }
            }
        }
        cp_wait<0>();
        __syncthreads();
    }
    float* C = p.C + (size_t)bz * p.sC;
    const int epi = p.epilogue;
    const float* extra = epi == GSR_EPI_ADD ? p.extra + (size_t)bz * p.s_extra : p.extra;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n = tid + h * T_THREADS;
        if (n >= N) continue;
        const float bias = epi == GSR_EPI_BIAS ? __ldg(extra + n) : 0.f;
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r) {
            if (r < rows) {
                float v = acc[r][h] * p.alpha + bias;
                if (epi == GSR_EPI_ADD) v += extra[(size_t)r * p.ld_extra + n];     // plain load: the addend may alias C
                C[(size_t)r * p.ldc + n] = v;
            }
        }
    }
    if (third && tid < 16 && 2 * T_THREADS + tid < N) {
        const int n = 2 * T_THREADS + tid;
        const float bias = epi == GSR_EPI_BIAS ? __ldg(extra + n) : 0.f;
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r) {
            if (r < rows) {
                float v = acc2[r] * p.alpha + bias;
                if (epi == GSR_EPI_ADD) v += extra[(size_t)r * p.ld_extra + n];
                C[(size_t)r * p.ldc + n] = v;
            }
        }
    }
}

// op(B) = B^T (B stored [n][k]): no staging of B at all.  Thread n streams its own row of B from global memory with
// 16-byte loads, eight in flight (one 128-byte line per thread), and op(A) comes from the same swizzled
// shared-memory copy; no barrier after the A staging.  108 us per 1024 subjects against 178 us on the tensor
// kernel (for row-major B the same scheme with coalesced 4-byte loads measured 90 us, the staged kernel 80 us).
constexpr int D_THREADS = 288;
template <bool TA, int R4>
__global__ void __launch_bounds__(D_THREADS, 4) tail_bt_kernel(const Params p) {
    __shared__ __align__(16) float As[T_KMAX * T_ROWS];
    const int tid = threadIdx.x;
    const int rows = p.M, K = p.K, N = p.N;
    const int bz = (int)gridDim.z - 1 - (int)blockIdx.z;       // reverse subject order: see tail_kernel
    const int kpad = (K + 31) & ~31;
    const float* A = p.A + (size_t)bz * p.sA;
    const float* B = p.B + (size_t)bz * p.sB;
    if (!TA) {
        float v[R4 * 4];
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r) v[r] = (r < rows && tid < K) ? __ldg(A + (size_t)r * p.lda + tid) : 0.f;
#pragma unroll
        for (int r = 0; r < R4 * 4; ++r)
            if (tid < kpad) As[tid * T_ROWS + ((((r >> 2) ^ (tid >> 1)) & 3) << 2) + (r & 3)] = v[r];
    } else {
        constexpr int PASSES = T_KMAX / (D_THREADS / 16);
        const int r = tid & 15;
        float v[PASSES];
#pragma unroll
        for (int j = 0; j < PASSES; ++j) {
            const int k = (tid >> 4) + j * (D_THREADS / 16);
            v[j] = (r < rows && k < K) ? __ldg(A + (size_t)k * p.lda + r) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < PASSES; ++j) {
            const int k = (tid >> 4) + j * (D_THREADS / 16);
            if (r < R4 * 4 && k < kpad) As[k * T_ROWS + ((((r >> 2) ^ (k >> 1)) & 3) << 2) + (r & 3)] = v[j];
        }
    }
    __syncthreads();
    const int n = tid;
    if (n >= N) return;
    float acc[R4 * 4];
#pragma unroll
    for (int r = 0; r < R4 * 4; ++r) acc[r] = 0.f;
    auto fma_k = [&](int k_static, const float* as, float b) {     // k_static: position inside the 32-row block
#pragma unroll
        for (int q = 0; q < R4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(as + k_static * T_ROWS + (((q ^ (k_static >> 1)) & 3) << 2));
            acc[4 * q + 0] = fmaf(a.x, b, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(a.y, b, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a.z, b, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a.w, b, acc[4 * q + 3]);
        }
    };
    const float4* b = reinterpret_cast<const float4*>(B + (size_t)n * p.ldb);
    for (int k0 = 0; k0 < kpad; k0 += 32) {
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = (k0 + 4 * i < K) ? __ldg(b + (k0 >> 2) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            fma_k(4 * i + 0, As + k0 * T_ROWS, bv[i].x); fma_k(4 * i + 1, As + k0 * T_ROWS, bv[i].y);
            fma_k(4 * i + 2, As + k0 * T_ROWS, bv[i].z); fma_k(4 * i + 3, As + k0 * T_ROWS, bv[i].w);
        }
    }
    float* C = p.C + (size_t)bz * p.sC;
    const int epi = p.epilogue;
    const float* extra = epi == GSR_EPI_ADD ? p.extra + (size_t)bz * p.s_extra : p.extra;
    const float bias = epi == GSR_EPI_BIAS ? __ldg(extra + n) : 0.f;
#pragma unroll
    for (int r = 0; r < R4 * 4; ++r) {
        if (r < rows) {
            float v = acc[r] * p.alpha + bias;
            if (epi == GSR_EPI_ADD) v += extra[(size_t)r * p.ld_extra + n];     // plain load: the addend may alias C
            C[(size_t)r * p.ldc + n] = v;
        }
    }
}

static bool tail_bt_ok(const Params& p, int transB) {
    return transB && p.M >= 1 && p.M <= T_ROWS && p.N <= D_THREADS && p.K <= T_KMAX && p.K % 4 == 0 && p.ldb % 4 == 0 && aligned16(p.B) && p.sB % 4 == 0;
}

static int launch_tail_bt(const Params& p, int transA, cudaStream_t st) {
    dim3 grid(1, 1, p.batch);
    const int r4 = (p.M + 3) / 4;
#define GSR_TAILBT_R4(TA)                                                                                        \
    do {                                                                                                          \
        if (r4 == 1) tail_bt_kernel<TA, 1><<<grid, D_THREADS, 0, st>>>(p);                                        \
        else if (r4 == 2) tail_bt_kernel<TA, 2><<<grid, D_THREADS, 0, st>>>(p);                                   \
        else if (r4 == 3) tail_bt_kernel<TA, 3><<<grid, D_THREADS, 0, st>>>(p);                                   \
        else tail_bt_kernel<TA, 4><<<grid, D_THREADS, 0, st>>>(p);                                                \
    } while (0)
    if (transA) GSR_TAILBT_R4(true); else GSR_TAILBT_R4(false);
#undef GSR_TAILBT_R4
    return launched("tcgemm_tail_bt");
}

static bool tail_ok(const Params& p, int transB) {     // p already points at the ragged rows
    return !transB && p.M >= 1 && p.M <= T_ROWS && p.N <= T_COLS && p.N % 4 == 0 && p.K <= T_KMAX &&
           p.ldb % 4 == 0 && aligned16(p.B) && p.sB % 4 == 0;
}

static int launch_tail(const Params& p, int transA, cudaStream_t st) {
    dim3 grid(1, 1, p.batch);
    const int r4 = (p.M + 3) / 4;
    constexpr int bytes = T_SMEM_FLOATS * (int)sizeof(float);
#define GSR_TAIL_LAUNCH(TA, R4)                                                                                  \
    do {                                                                                                          \
        auto k = tail_kernel<TA, R4>;                                                                             \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);              \
        if (e != cudaSuccess) return fail((int)e, "tcgemm tail: cudaFuncSetAttribute failed");                   \
        k<<<grid, T_THREADS, bytes, st>>>(p);                                                                     \
    } while (0)
#define GSR_TAIL_R4(TA)                                                                                          \
    do {                                                                                                          \
        if (r4 == 1) GSR_TAIL_LAUNCH(TA, 1); else if (r4 == 2) GSR_TAIL_LAUNCH(TA, 2);                            \
        else if (r4 == 3) GSR_TAIL_LAUNCH(TA, 3); else GSR_TAIL_LAUNCH(TA, 4);                                    \
    } while (0)
    if (transA) GSR_TAIL_R4(true); else GSR_TAIL_R4(false);
#undef GSR_TAIL_R4
#undef GSR_TAIL_LAUNCH
    return launched("tcgemm_tail");
}
}  // namespace tail

// C[m][n] = alpha * sum_z ws[z][m][n] (+ extra[m][n]); fixed summation order.  Rows below m_head hold
// `splits` partials of pitch M * N; the rows from m_head on (the ragged tile after the CTA pairs, launched on its
// own) hold `splits_tail` partials of pitch (M - m_head) * N at ws_tail.
__global__ void reduce_partials_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                       int splits, int m_head, const float* __restrict__ ws_tail, int splits_tail,
                                       float alpha, const float* __restrict__ extra, int ld_extra, int m_pitch = 0) {
    // rows [0, M) are written; the head partials have a pitch of m_pitch * N (m_pitch = M unless fewer rows are reduced here)
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * N) return;
    const int m = e / N, n = e - m * N;
    if (m_pitch == 0) m_pitch = M;
    float acc = 0.f;
    if (m < m_head) {
        for (int z = 0; z < splits; ++z) acc += __ldg(ws + (size_t)z * m_pitch * N + e);
    } else {
        const size_t pitch = (size_t)(m_pitch - m_head) * N;
        const float* src = ws_tail + (size_t)(m - m_head) * N + n;
        for (int z = 0; z < splits_tail; ++z) acc += __ldg(src + z * pitch);
    }
    acc *= alpha;
    if (extra) acc += __ldg(extra + (size_t)m * ld_extra + n);
    C[(size_t)m * ldc + n] = acc;
}

// C[m_head + r][n] = alpha * sum_b T[b][r][n] (+ extra): the ragged rows of a subject-reduced product, computed per subject
// by the FP32 ragged-row kernels, summed over the subjects in a fixed order (8 phases of b, then the phases).
__global__ void __launch_bounds__(256) batch_sum_rows_kernel(const float* __restrict__ T, float* __restrict__ C, int rows, int N, int ldc,
                                                             int m_head, int batch, float alpha, const float* __restrict__ extra, int ld_extra) {
    __shared__ float red[8][33];
    const int e = blockIdx.x * 32 + threadIdx.x, total = rows * N;
    float acc = 0.f;
    if (e < total)
        for (int b = threadIdx.y; b < batch; b += 8) acc += __ldg(T + (size_t)b * total + e);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && e < total) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += red[j][threadIdx.x];
        const int r = e / N, n = e - r * N;
        sum *= alpha;
        if (extra) sum += __ldg(extra + (size_t)(m_head + r) * ld_extra + n);
        C[(size_t)(m_head + r) * ldc + n] = sum;
    }
}

// Split counts of gsr_tcgemm_reduce_f32: `splits` CTAs per tile on the one-CTA path (every SM busy once);
// on the CTA-pair path the pairs share the SMs among themselves and the ragged tile is launched on its own.
struct ReducePlan { int pairs, m_head, pair_splits, tail_splits; size_t floats; };
static ReducePlan reduce_plan(int M, int N, int batch, int splits) {
    ReducePlan r{0, M, splits, 0, (size_t)splits * M * N};
    const int tiles_m = (M + BM - 1) / BM;
    if (!g_pairs || tiles_m < 2 || N > BN) return r;
    r.pairs = tiles_m / 2;
    r.m_head = std::min(M, r.pairs * 2 * BM);
    r.pair_splits = std::max(1, std::min(batch, kNumSMs / (2 * r.pairs)));
    if (r.m_head < M) r.tail_splits = std::max(1, std::min(batch, kNumSMs));
    r.floats = std::max(r.floats, (size_t)r.pair_splits * M * N + (size_t)r.tail_splits * (M - r.m_head) * N);
    // the TMA route (csrc/tcgemm_tma.cu): 256-row head tiles, <= 16 ragged rows per subject on the FP32 kernels
    const int head_tiles = tma_head_tiles(M);
    if (head_tiles > 0) {
        const int m_head = tma_head_rows(M), ps = std::max(1, std::min(batch, kNumSMs / (2 * head_tiles)));
        r.floats = std::max(r.floats, (size_t)ps * M * N + (size_t)std::max(batch, kNumSMs) * (M - m_head) * N);
    }
    return r;
}

}  // namespace tc
}  // namespace gsr

using namespace gsr;

#ifdef GSR_TC_TRACE
extern "C" __attribute__((visibility("default"))) int gsr_tcgemm_trace_read(unsigned long long* host, int count) {
    return (int)cudaMemcpyFromSymbol(host, tc::g_trace, sizeof(unsigned long long) * count);
}
#endif

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
                 epilogue, extra, ld_extra, stride_extra};
    if (!tc::operands_ok(p)) return fail(GSR_E_ALIGN, "tcgemm: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    if (tc::g_pairs && tc::tma_ok(p, transA, transB)) {
        // TMA-fed persistent CTA pairs (csrc/tcgemm_tma.cu) for the first 256 floor(M / 256) rows; the remaining rows on the
        // FP32 ragged-row kernels (<= 16 rows: the 12 rows a 268-row subject matrix leaves) or the one-CTA tensor kernel
        int rc = tc::tma_launch(p, transA, transB, false, batch, (cudaStream_t)stream);
        if (rc) return rc;
        const int m_done = tc::tma_head_rows(M);
        if (m_done == M) return 0;
        tc::Params t = p;
        t.M = M - m_done;
        t.A = A + (transA ? (size_t)m_done : (size_t)m_done * lda);
        t.C = C + (size_t)m_done * ldc;
        if (epilogue == GSR_EPI_ADD) t.extra = extra + (size_t)m_done * ld_extra;
        if (tc::g_tail && tc::tail::tail_bt_ok(t, transB)) return tc::tail::launch_tail_bt(t, transA, (cudaStream_t)stream);
        if (tc::g_tail && tc::tail::tail_ok(t, transB)) return tc::tail::launch_tail(t, transA, (cudaStream_t)stream);
        dim3 gt((t.M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN, batch);
        return tc::launch<false>(t, transA, transB, gt, (cudaStream_t)stream);
    }
    if (tc::g_pairs && tc::pair_ok(p) && (long long)(transA ? K : M) * lda < 0x7fffffffLL && (long long)(transB ? N : K) * ldb < 0x7fffffffLL) {
        // CTA pairs (cta_group::2) for the first 2 * floor(tiles / 2) M tiles, the one-CTA kernel for an odd last tile
        int rc = tc::launch_pairs<false>(p, transA, transB, batch, (cudaStream_t)stream);
        if (rc) return rc;
        const int tiles_m = (M + tc::BM - 1) / tc::BM;
        if (tiles_m % 2 == 0) return 0;
        const int m_done = (tiles_m - 1) * tc::BM;
        tc::Params t = p;
        t.M = M - m_done;
        t.A = A + (transA ? (size_t)m_done : (size_t)m_done * lda);
        t.C = C + (size_t)m_done * ldc;
        if (epilogue == GSR_EPI_ADD) t.extra = extra + (size_t)m_done * ld_extra;
        if (tc::g_tail && tc::tail::tail_bt_ok(t, transB)) return tc::tail::launch_tail_bt(t, transA, (cudaStream_t)stream);
        if (tc::g_tail && tc::tail::tail_ok(t, transB)) return tc::tail::launch_tail(t, transA, (cudaStream_t)stream);
        dim3 gt(1, (N + tc::BN - 1) / tc::BN, batch);
        return tc::launch<false>(t, transA, transB, gt, (cudaStream_t)stream);
    }
    dim3 grid((M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN, batch);
    if (grid.y > 65535) return fail(GSR_E_UNSUPPORTED, "tcgemm: N too large");
    return tc::launch<false>(p, transA, transB, grid, (cudaStream_t)stream);
}

// C = alpha (op(A1) op(B1) + op(A2) op(B2)): two products of the same shape and operand majors into one result -- the two
// contributions to the gradient of the adjacency (models/layers.py:85-86 applied twice, models/gsr_model.py:30-32).  On the
// TMA kernel both accumulate in tensor memory (one k-loop over the k-blocks of the first pair, then the second): no addend is
// read back, one result is written.  The <= 16 ragged rows run on the FP32 kernels as two launches, the second adding in place.
extern "C" int gsr_tcgemm_dual_f32(const float* A1, const float* B1, const float* A2, const float* B2, float* C,
                                   int M, int N, int K1, int K2, int lda1, int ldb1, int lda2, int ldb2, int ldc,
                                   long long strideA1, long long strideB1, long long strideA2, long long strideB2, long long strideC,
                                   int transA, int transB, int batch, float alpha, void* stream) {
    if (!A1 || !B1 || !A2 || !B2 || !C || M < 0 || N < 0 || K1 < 1 || K2 < 1 || batch < 0) return fail(GSR_E_ARG, "tcgemm_dual: bad argument");
    if (M == 0 || N == 0 || batch == 0) return 0;
    if (batch > 65535) return fail(GSR_E_UNSUPPORTED, "tcgemm_dual: batch > 65535");
    tc::Params p{A1, B1, C, M, N, K1, lda1, ldb1, ldc, strideA1, strideB1, strideC, batch, 1, alpha, GSR_EPI_NONE, nullptr, 0, 0};
    tc::Params q{A2, B2, C, M, N, K2, lda2, ldb2, ldc, strideA2, strideB2, strideC, batch, 1, alpha, GSR_EPI_NONE, nullptr, 0, 0};
    if (!tc::operands_ok(p) || !tc::operands_ok(q)) return fail(GSR_E_ALIGN, "tcgemm_dual: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    const bool same_batching = (strideA1 != 0) == (strideA2 != 0) && (strideB1 != 0) == (strideB2 != 0);
    if (tc::g_pairs && tc::g_dual && same_batching && tc::tma_ok(p, transA, transB) && tc::tma_ok(q, transA, transB)) {
        int rc = tc::tma_launch(p, transA, transB, false, batch, (cudaStream_t)stream, &q);
        if (rc) return rc;
        const int m_done = tc::tma_head_rows(M);
        if (m_done == M) return 0;
        for (int pass = 0; pass < 2; ++pass) {
            tc::Params t = pass ? q : p;
            t.M = M - m_done;
            t.A = t.A + (transA ? (size_t)m_done : (size_t)m_done * t.lda);
            t.C = C + (size_t)m_done * ldc;
            if (pass) { t.epilogue = GSR_EPI_ADD; t.extra = t.C; t.ld_extra = ldc; t.s_extra = strideC; }
            if (tc::g_tail && tc::tail::tail_bt_ok(t, transB)) rc = tc::tail::launch_tail_bt(t, transA, (cudaStream_t)stream);
            else if (tc::g_tail && tc::tail::tail_ok(t, transB)) rc = tc::tail::launch_tail(t, transA, (cudaStream_t)stream);
            else rc = tc::launch<false>(t, transA, transB, dim3((t.M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN, batch), (cudaStream_t)stream);
            if (rc) return rc;
        }
        return 0;
    }
    // anything else: two launches, the second with the first one's result as its addend
    int rc = gsr_tcgemm_f32(A1, B1, C, M, N, K1, lda1, ldb1, ldc, strideA1, strideB1, strideC, transA, transB, batch, alpha,
                            GSR_EPI_NONE, nullptr, 0, 0, stream);
    if (rc) return rc;
    return gsr_tcgemm_f32(A2, B2, C, M, N, K2, lda2, ldb2, ldc, strideA2, strideB2, strideC, transA, transB, batch, alpha,
                          GSR_EPI_ADD, C, ldc, strideC, stream);
}

extern "C" size_t gsr_tcgemm_reduce_workspace(int M, int N, int batch, int* splits_out) {
    const int tiles = ((N + tc::BN - 1) / tc::BN) * ((M + tc::BM - 1) / tc::BM);
    int splits = (kNumSMs + tiles - 1) / tiles;          // one CTA per SM (TMEM is allocated whole)
    if (splits > batch) splits = batch;
    if (splits < 1) splits = 1;
    if (splits_out) *splits_out = splits;
    return tc::reduce_plan(M, N, batch, splits).floats;
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
    tc::Params p{A, B, workspace, M, N, K, lda, ldb, N, strideA, strideB, (long long)M * N, batch, splits, 1.0f,
                 GSR_EPI_NONE, nullptr, 0, 0};
    if (!tc::operands_ok(p)) return fail(GSR_E_ALIGN, "tcgemm_reduce: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    const int tiles_m = (M + tc::BM - 1) / tc::BM;
    const tc::ReducePlan plan = tc::reduce_plan(M, N, batch, splits);
    const float* ws_tail = workspace;
    int rc;
    if (tc::g_pairs && N <= 272 && tc::tma_ok(p, transA, transB)) {
        // TMA route: persistent CTA pairs accumulate the subjects of a split in tensor memory for the 256-row head tiles
        // (a last tile of >= 64 rows included: the whole of a 160-row product); <= 16 ragged rows (the 12 rows a 268-row
        // product leaves) are computed per subject by the FP32 ragged-row kernels and summed over the subjects
        const int head_tiles = tc::tma_head_tiles(M), m_head = tc::tma_head_rows(M);
        const int ps = std::max(1, std::min(batch, kNumSMs / (2 * head_tiles)));
        p.splits = ps;
        rc = tc::tma_launch(p, transA, transB, true, ps, (cudaStream_t)stream);
        if (rc) return rc;
        int tail_splits = 0;
        float* ws_t = workspace + (size_t)ps * M * N;
        if (m_head < M) {
            tc::Params q = p;
            q.A = transA ? A + m_head : A + (size_t)m_head * lda;
            q.C = ws_t;
            q.M = M - m_head;
            q.ldc = N;
            q.sC = (long long)q.M * N;
            q.alpha = 1.0f; q.epilogue = GSR_EPI_NONE; q.extra = nullptr;
            const bool simt = tc::g_tail && tc::g_reduce_simt_tails && (tc::tail::tail_bt_ok(q, transB) || tc::tail::tail_ok(q, transB));
            if (simt) {
                rc = tc::tail::tail_bt_ok(q, transB) ? tc::tail::launch_tail_bt(q, transA, (cudaStream_t)stream)
                                                     : tc::tail::launch_tail(q, transA, (cudaStream_t)stream);
                if (rc) return rc;
                const int total_t = q.M * N;
                tc::batch_sum_rows_kernel<<<(total_t + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(
                    ws_t, C, q.M, N, ldc, m_head, batch, alpha, epilogue == GSR_EPI_ADD ? extra : nullptr, ld_extra);
                if ((rc = launched("tcgemm_reduce_tail_sum"))) return rc;
                tail_splits = -1;                                    // rows m_head.. are final
            } else {
                q.splits = tail_splits = std::max(1, std::min(batch, kNumSMs));
                rc = tc::launch<true>(q, transA, transB, dim3((q.M + tc::BM - 1) / tc::BM, 1, q.splits), (cudaStream_t)stream);
                if (rc) return rc;
            }
        }
        const int rows_out = tail_splits < 0 ? m_head : M;           // reduce_partials_kernel covers rows [0, rows_out)
        tc::reduce_partials_kernel<<<(rows_out * N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            workspace, C, rows_out, N, ldc, ps, m_head, ws_t, tail_splits < 0 ? 0 : tail_splits, alpha,
            epilogue == GSR_EPI_ADD ? extra : nullptr, ld_extra, M);
        return launched("tcgemm_reduce_partials");
    }
    if (plan.pairs && tc::pair_ok(p)) {
        // CTA pairs take the even number of leading M tiles; the ragged last tile is launched on its own with its
        // own split count and compact partials behind those of the pairs.
        p.splits = plan.pair_splits;
        if (tc::tma_ok(p, transA, transB) && (M / 256) * 256 == plan.m_head) rc = tc::tma_launch(p, transA, transB, true, plan.pair_splits, (cudaStream_t)stream);
        else rc = tc::launch_pairs<true>(p, transA, transB, plan.pair_splits, (cudaStream_t)stream);
        if (rc) return rc;
        if (plan.m_head < M) {
            tc::Params q = p;
            q.splits = plan.tail_splits;
            q.A = transA ? A + plan.m_head : A + (size_t)plan.m_head * lda;
            q.C = workspace + (size_t)plan.pair_splits * M * N;
            q.M = M - plan.m_head;
            q.sC = (long long)q.M * N;
            ws_tail = q.C;
            rc = tc::launch<true>(q, transA, transB, dim3(1, 1, q.splits), (cudaStream_t)stream);
            if (rc) return rc;
        }
        splits = plan.pair_splits;
    } else {
        dim3 grid(tiles_m, (N + tc::BN - 1) / tc::BN, splits);
        rc = tc::launch<true>(p, transA, transB, grid, (cudaStream_t)stream);
        if (rc) return rc;
    }
    const int total = M * N;
    tc::reduce_partials_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        workspace, C, M, N, ldc, splits, (plan.pairs && tc::pair_ok(p)) ? plan.m_head : M, ws_tail, plan.tail_splits, alpha,
        epilogue == GSR_EPI_ADD ? extra : nullptr, ld_extra);
    return launched("tcgemm_reduce_partials");
}
