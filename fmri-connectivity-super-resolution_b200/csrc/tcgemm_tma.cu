// TMA-fed persistent CTA-pair GEMM (tcgen05.mma.cta_group::2 kind::tf32, 3xTF32 split) -- round 2.
//
// Same arithmetic as csrc/tcgemm.cu (lo.hi + hi.lo + hi.hi into one fp32 accumulator in tensor memory), new data path:
//   * operands arrive by `cp.async.bulk.tensor` (TMA, 3-D tensor maps {columns, rows, batch item}: out-of-range
//     rows / columns of a subject matrix are zero filled per subject) straight into the shared-memory layouts the
//     UMMA descriptors name -- K-major operands with CU_TENSOR_MAP_SWIZZLE_128B (box 32 floats x rows; descriptor
//     layout SWIZZLE_128B, SBO 1024), MN-major operands with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (one box of
//     32 floats x 32 k-rows per 32-wide block; layout SWIZZLE_128B_BASE32B, LBO 4096, SBO 512).  Probed on the
//     hardware first: tools/probes/tma_probe.cu, profiles/r2_tma_tf32_probe.txt;
//   * kind::tf32 reads 32-bit containers and IGNORES the 13 low mantissa bits (same probe, T1: 3072 of 3072
//     full-mantissa values were read truncated), so the raw fp32 tile IS the `hi` term; 8 converter warps only
//     compute lo = tf32(x - trunc(x)) shared memory -> shared memory (element-wise, layout agnostic, conflict free):
//     no global-load scoreboard, no register staging, half the stores of the thread-staged kernels;
//   * persistent: one CTA pair per TPC walks the tiles (256 rows x all columns each); the TMA warp runs up to three
//     k-blocks ahead, across tile boundaries, so the pipeline fill and the epilogue of a tile overlap the loads of
//     the next one;
//   * epilogue: 4 warps drain the accumulator (tcgen05.ld 32x32b.x32), apply alpha / bias / addend, stage 32 x 32
//     blocks in a 128B-swizzled double buffer and write them with TMA stores (clipped at the matrix edge).
// Warp roles (448 threads): 0 TMA producer, 1 MMA issuer (leader CTA) + tensor-memory allocation, 2-9 converters,
// 10-13 epilogue (warp % 4 = tensor-memory lane quarter).
#include <cuda.h>

#include "tcgemm.cuh"

namespace gsr {
namespace tc {

int g_tma = 1;
int g_utccp = 0;            // gsr_set_tuning("tcgemm_utccp", 0|1): K-major A through tcgen05.cp + tensor memory / from shared memory (default).
                            // Measured (profiles/r2_tcgemm_tma_notes.txt): exact, and it removes 66 KB of the 252 KB of shared-memory
                            // traffic per k-block, but the copies execute IN the tensor pipe, in order with the MMAs (8 x ~64 clk per
                            // k-block): the MMA thread's issue -> commit time goes from 1.15 to 1.5 us and the stacked product from
                            // 0.259 to 0.277 ms.  Kept as a switch.
int g_tmem_a = 0;           // gsr_set_tuning("tcgemm_tmem_a", 0|1): the hi term of op(A) in TENSOR MEMORY, written by the converter warps
                            // with tcgen05.st while they compute lo -- four of the six MMAs of a k-step then read A there instead of
                            // re-reading 4 KB of shared memory each (66 of the 252 KB of shared-memory traffic per k-block).  Exact, but
                            // measured no faster (stacked NT 0.265 against 0.259 ms): the MMAs of a k-block take 1.15-1.25 us with either
                            // operand source (the tensor pipe, not its shared-memory reads, is what they wait for), and the two 32-column
                            // A stages leave room for a 4-chunk instead of a 2-chunk accumulator overlap.  Off by default.
int g_partial_head = 1;     // a last 256-row tile with >= 64 valid rows runs on the TMA kernel (the 160-row subject-reduced product)
int g_reduce_simt_tails = 1;    // <= 16 ragged rows of a subject-reduced product: FP32 ragged-row kernels per subject + a sum over subjects
int g_cluster_release = 0;  // gsr_set_tuning("tcgemm_cluster_release", 0|1): see mbar_arrive_remote (csrc/tcgemm.cuh)
int g_pf_ahead = 0;         // gsr_set_tuning("tcgemm_prefetch", k): k-blocks of L2 prefetch (cp.async.bulk.prefetch.tensor) in front of the
                            // TMA loads; default off -- measured monotonically slower (stacked NN 0.272 / 0.290 / 0.291 / 0.287 / 0.310 ms
                            // for k = 0 / 3 / 6 / 9 / 12): the load latency is not what bounds the k-block rate
int g_mn_partial = 1;       // gsr_set_tuning("tcgemm_mn_partial", 0|1): MN-major B ranges may end inside a 32-column block (256 + 16)

namespace t3 {

#ifdef GSR_T3_TRACE
// Pipeline trace (tools/tc3_trace.py builds a private copy of the library with -DGSR_T3_TRACE): cluster 1 stamps
// %globaltimer at its pipeline events; slot layout [rank][role][k-block or tile][event].
constexpr int TR_ITS = 64, TR_EVS = 12;
__device__ unsigned long long g_trace[2 * 4 * TR_ITS * TR_EVS];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define T3_STAMP(role, it, ev) do { if ((blockIdx.x >> 1) == 1 && (it) < TR_ITS) g_trace[(((blockIdx.x & 1) * 4 + (role)) * TR_ITS + (it)) * TR_EVS + (ev)] = gtime(); } while (0)
#else
#define T3_STAMP(role, it, ev) do { } while (0)
#endif

constexpr int RAW_STAGES = 3, LO_STAGES = 2;
constexpr int A_TILE = BM * BK * 4;                       // 16384
constexpr int B_TILE_MAX = 160 * BK * 4;                  // 20480 (MN-major: 96 + 64 columns; K-major: 80 + 56 rows = 17408)
constexpr int SLOT_BYTES = A_TILE + B_TILE_MAX;           // 36864, multiple of 1024
constexpr int CONV_WARPS = 8, EPI_WARPS = 4;
constexpr int THREADS = 32 * (2 + CONV_WARPS + EPI_WARPS);        // 448
constexpr int STG_BYTES = 32 * 32 * 4;                    // one 32 x 32 staging block
constexpr int SMEM_DATA = (RAW_STAGES + LO_STAGES) * SLOT_BYTES + EPI_WARPS * 2 * STG_BYTES;      // 217088
constexpr int SMEM_BYTES = SMEM_DATA + 256 + 1024;        // + barriers / tensor-memory slot, + alignment slack
constexpr int TMEM_COLS = 512;

struct Plan {
    int tiles_m;            // pair tiles along M per batch item
    int tiles_n, n_tile;    // column tiles per row tile (1 unless N > 272: then 256 columns each, MLP-sized products)
    int items;              // output batch items (REDUCE: splits)
    int nsub;               // REDUCE: subjects accumulated in all (else 1 per item)
    int splits;             // REDUCE: stride between the subjects of an item
    int M, N, K, nkb;
    int K2, nkb1;           // dual product C = op(A) op(B) + op(A2) op(B2): k-blocks [0, nkb1) come from the first pair of operands
                            // (K columns), [nkb1, nkb) from the second (K2 columns); nkb1 = nkb when there is one pair
    int nr0, nr1;           // UMMA N of the two accumulator column ranges (nr1 may be 0)
    int acc_base1;          // tensor-memory column of the odd tiles' accumulator (even tiles: 0); the two overlap when 2 (nr0 + nr1) > 512
    int a_ring;             // tensor-memory column of the A-operand ring (tcgen05.cp, K-major A only); 0 = A is read from shared memory
    int a_hi_col;           // tensor-memory column of the A-hi stages (2 x 32 columns, written by the converters); 0 = off
    int b_off1;             // byte offset of the second range inside a B tile
    int b_bytes;            // bytes of this CTA's B tile per k-block (TMA transaction count)
    int pf_ahead;           // k-blocks the L2 prefetch cursor runs in front of the loads (0 = off)
    int cluster_release;    // remote mbarrier arrivals with release.cluster (1) or the default release.cta (0)
    int a_batched, b_batched;
    float alpha;
    int epilogue;
    const float* extra; int ld_extra; long long s_extra;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" :: "l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory"); }

// smem -> tensor memory copy of a 128-lane x 256-bit slice (8 fp32 columns) in both CTAs of the pair
__device__ __forceinline__ void utccp2_128x256b(uint32_t taddr, uint32_t d_lo, uint32_t d_hi) {
    asm volatile("{\n\t.reg .b64 da;\n\tmov.b64 da, {%1, %2};\n\ttcgen05.cp.cta_group::2.128x256b [%0], da;\n\t}"
                 :: "r"(taddr), "r"(d_lo), "r"(d_hi) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma2_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// registers -> tensor memory: lane = this thread's row, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.ld without the wait, and a wait the compiler cannot move uses of the destination registers across
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// TA: op(A) = A^T (A stored [K][M], MN-major).  TB: op(B) = B^T (B stored [N][K], K-major).
template <bool TA, bool TB, bool REDUCE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tcgemm3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapB1, const __grid_constant__ CUtensorMap mapC,
               const __grid_constant__ CUtensorMap mapE, const __grid_constant__ CUtensorMap mapA2,
               const __grid_constant__ CUtensorMap mapB02, const __grid_constant__ CUtensorMap mapB12, const Plan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* raw = smem;                                              // [RAW_STAGES][A | B]  fp32 as landed = hi
    uint8_t* lo = smem + RAW_STAGES * SLOT_BYTES;                     // [LO_STAGES][A | B]   tf32(x - trunc(x))
    uint8_t* stg = lo + LO_STAGES * SLOT_BYTES;                       // [EPI_WARPS][2][32 x 32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
    // raw_full[3] (TMA bytes landed, local), raw_empty[3] / lo_empty[2] (MMA retired: commit multicast),
    // conv_full[2] (leader's: converter warps of both CTAs), tmem_full (commit multicast), tmem_empty (leader's)
    const uint32_t bar_raw_full = smem_u32(bars), bar_raw_empty = smem_u32(bars + RAW_STAGES);
    const uint32_t bar_lo_empty = smem_u32(bars + 2 * RAW_STAGES), bar_conv_full = smem_u32(bars + 2 * RAW_STAGES + LO_STAGES);
    const uint32_t bar_tmem_full = smem_u32(bars + 2 * RAW_STAGES + 2 * LO_STAGES), bar_tmem_empty = bar_tmem_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RAW_STAGES + 2 * LO_STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int tiles_mn = p.tiles_m * p.tiles_n;                       // tile = (item z, row tile mp, column tile nt), nt fastest
    const int total_tiles = p.items * tiles_mn;
    const int R0 = p.nr0 >> 1, R1 = p.nr1 >> 1;                       // this CTA's share of the two B ranges
    const uint32_t b_bytes = (uint32_t)p.b_bytes;
    const int kb_per_tile = (REDUCE ? (p.nsub + p.splits - 1) / p.splits : 1) * p.nkb;   // upper bound, exact below

    if (tid == 0) {
        for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(bar_raw_full + 8 * s, 1); mbar_init(bar_raw_empty + 8 * s, 1); }
        for (int s = 0; s < LO_STAGES; ++s) { mbar_init(bar_lo_empty + 8 * s, 1); mbar_init(bar_conv_full + 8 * s, 2 * CONV_WARPS); }
        mbar_init(bar_tmem_full, 1);
        mbar_init(bar_tmem_empty, 2 * EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
    if (warp == 0 && lane == 0) { prefetch_map(&mapA); prefetch_map(&mapB0); prefetch_map(&mapB1); prefetch_map(&mapC); prefetch_map(&mapE); prefetch_map(&mapA2); prefetch_map(&mapB02); prefetch_map(&mapB12); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    (void)kb_per_tile;

    if (warp == 0) {
        // ===== TMA producer (one lane): every CTA loads its own 128 rows of op(A) and its half of op(B).
        // An optional second cursor runs p.pf_ahead k-blocks in front of the loads and issues the same boxes as
        // cp.async.bulk.prefetch.tensor (L2 only); measured slower, off by default (g_pf_ahead).
        if (lane == 0) {
            struct Blk { int tile, sub, kb; };
            auto nsub_of = [&](int tile) { return REDUCE ? (p.nsub - tile / tiles_mn + p.splits - 1) / p.splits : 1; };
            auto next = [&](Blk& b) {
                if (++b.kb == p.nkb) { b.kb = 0; if (++b.sub == nsub_of(b.tile)) { b.sub = 0; b.tile += n_clusters; } }
            };
            // issue the boxes of one k-block: PREFETCH -> into L2, else into raw slot `s` signalling `full`
            auto boxes = [&](const Blk& blk, bool prefetch, uint32_t a_dst, uint32_t full) {
                const int z = blk.tile / tiles_mn, mn = blk.tile - z * tiles_mn, mp = mn / p.tiles_n;
                const int m0 = mp * 256 + (int)rank * BM, n0 = (mn - mp * p.tiles_n) * p.n_tile;
                const int b = REDUCE ? z + blk.sub * p.splits : z;
                const int ba = p.a_batched ? b : 0, bb = p.b_batched ? b : 0;
                const bool second = blk.kb >= p.nkb1;                                               // dual product: the second pair of operands
                const int k0 = (second ? blk.kb - p.nkb1 : blk.kb) * BK;
                const CUtensorMap* mA = second ? &mapA2 : &mapA;
                const CUtensorMap* mB0 = second ? &mapB02 : &mapB0;
                const CUtensorMap* mB1 = second ? &mapB12 : &mapB1;
                const uint32_t b_dst = a_dst + A_TILE;
                auto ld = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2) {
                    if (prefetch) tma_prefetch(map, c0, c1, c2); else tma_load(dst, map, full, c0, c1, c2);
                };
                if (!TA) ld(a_dst, mA, k0, m0, ba);                                                 // box {32 k, 128 m}
                else {
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) ld(a_dst + j * 4096, mA, m0 + 32 * j, k0, ba);           // box {32 m, 32 k}
                }
                if (TB) {                                                                           // boxes {32 k, R0 n}, {32 k, R1 n}
                    ld(b_dst, mB0, k0, n0 + R0 * (int)rank, bb);
                    if (R1 > 0) ld(b_dst + p.b_off1, mB1, k0, n0 + p.nr0 + R1 * (int)rank, bb);
                } else {                                                                            // box {32 n, 32 k} per 32 columns
                    // (a range whose share is not a multiple of 32 columns uses the head of its last block only)
                    for (int j = 0; j < (R0 + 31) / 32; ++j) ld(b_dst + j * 4096, mB0, n0 + R0 * (int)rank + 32 * j, k0, bb);
                    for (int j = 0; j < (R1 + 31) / 32; ++j) ld(b_dst + p.b_off1 + j * 4096, mB0, n0 + p.nr0 + R1 * (int)rank + 32 * j, k0, bb);
                }
            };
            Blk cur{cluster_id, 0, 0}, pf{cluster_id, 0, 0};
            const int pf_ahead = p.pf_ahead;
            for (int i = 0; i < pf_ahead && pf.tile < total_tiles; ++i) { boxes(pf, true, 0, 0); next(pf); }
            for (uint32_t it = 0; cur.tile < total_tiles; ++it, next(cur)) {
                if (pf_ahead > 0 && pf.tile < total_tiles) { boxes(pf, true, 0, 0); next(pf); }
                const uint32_t s = it % RAW_STAGES, use = it / RAW_STAGES;
                T3_STAMP(0, it, 0);
                if (use > 0) mbar_wait(bar_raw_empty + 8 * s, (use - 1) & 1);
                T3_STAMP(0, it, 1);
                const uint32_t full = bar_raw_full + 8 * s;
                mbar_expect_tx(full, (uint32_t)A_TILE + b_bytes);
                boxes(cur, false, smem_u32(raw + s * SLOT_BYTES), full);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one lane of the leader CTA
        if (rank == 0 && lane == 0) {
            const uint32_t idesc0 = instr_desc2(p.nr0, TA ? 1 : 0, TB ? 0 : 1);
            const uint32_t idesc1 = instr_desc2(p.nr1 > 0 ? p.nr1 : 16, TA ? 1 : 0, TB ? 0 : 1);
            // A in tensor memory is always "K-major" (lane = row, column = k)
            const uint32_t idesc0t = instr_desc2(p.nr0, 0, TB ? 0 : 1), idesc1t = instr_desc2(p.nr1 > 0 ? p.nr1 : 16, 0, TB ? 0 : 1);
            // K-major (SWIZZLE_128B): 8-row groups 1024 bytes apart (SBO), k-step of 8 floats = +32 bytes
            // MN-major (SWIZZLE_128B_BASE32B): 32-wide blocks 4096 bytes apart (LBO), 4-k groups 512 bytes apart (SBO), k-step = +1024
            const uint32_t a_dhi = TA ? desc_hi(512, 1) : desc_hi(1024, 2), b_dhi = TB ? desc_hi(1024, 2) : desc_hi(512, 1);
            const uint32_t a_lbo = TA ? 4096 : 16, b_lbo = TB ? 16 : 4096;
            const uint32_t a_step = (TA ? 1024u : 32u) >> 4, b_step = (TB ? 32u : 1024u) >> 4;
            const uint32_t n1 = (uint32_t)p.b_off1 >> 4;                    // start of the second range inside a B tile
            uint32_t it = 0, tile_n = 0, kstep_n = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tile_n) {
                const int z = tile / tiles_mn;
                const int nsub = REDUCE ? (p.nsub - z + p.splits - 1) / p.splits : 1;
                T3_STAMP(1, it, 3);
                // every epilogue warp of both CTAs has read the columns tile t - 1 shares with this tile's accumulator
                // (and, before that, all of tile t - 2, which used the same columns as this one)
                if (tile_n > 0) { mbar_wait_cluster(bar_tmem_empty, (tile_n - 1) & 1); tc_fence_after(); }
                T3_STAMP(1, it, 4);
                // even tiles accumulate at column 0, odd tiles at acc_base1: the drain of tile t overlaps the MMAs of tile
                // t + 1 (the epilogue releases the accumulator as soon as the columns the two share have been read)
                const uint32_t tacc = tmem_base + ((tile_n & 1) ? (uint32_t)p.acc_base1 : 0u);
                uint32_t acc = 0;
                for (int sub = 0; sub < nsub; ++sub) {
                    for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                        const uint32_t rs = it % RAW_STAGES, ls = it % LO_STAGES, luse = it / LO_STAGES;
                        T3_STAMP(1, it, 0);
                        mbar_wait_cluster(bar_conv_full + 8 * ls, luse & 1);
                        tc_fence_after();
                        T3_STAMP(1, it, 1);
                        const uint32_t hb = smem_u32(raw + rs * SLOT_BYTES), lb = smem_u32(lo + ls * SLOT_BYTES);
                        uint32_t ah = desc_lo(hb, a_lbo), al = desc_lo(lb, a_lbo);
                        uint32_t bh = desc_lo(hb + A_TILE, b_lbo), bl = desc_lo(lb + A_TILE, b_lbo);
                        const int ksteps = kb < p.nkb1 ? min(BK / 8, (p.K - kb * BK + 7) >> 3) : min(BK / 8, (p.K2 - (kb - p.nkb1) * BK + 7) >> 3);
                        if (p.a_hi_col) {
                            // A-hi in tensor memory (converter warps, tcgen05.st): lo.hi reads A-lo from shared memory,
                            // hi.lo and hi.hi read A-hi at 8 columns per k-step of stage `ls`
                            const uint32_t t_a = tmem_base + (uint32_t)p.a_hi_col + 32u * ls;
                            for (int j = 0; j < ksteps; ++j) {
                                umma2_tf32(tacc, al, a_dhi, bh, b_dhi, idesc0, acc);
                                umma2_tf32_ts(tacc, t_a + 8u * j, bl, b_dhi, idesc0t, 1u);
                                umma2_tf32_ts(tacc, t_a + 8u * j, bh, b_dhi, idesc0t, 1u);
                                if (p.nr1 > 0) {
                                    umma2_tf32(tacc + p.nr0, al, a_dhi, bh + n1, b_dhi, idesc1, acc);
                                    umma2_tf32_ts(tacc + p.nr0, t_a + 8u * j, bl + n1, b_dhi, idesc1t, 1u);
                                    umma2_tf32_ts(tacc + p.nr0, t_a + 8u * j, bh + n1, b_dhi, idesc1t, 1u);
                                }
                                acc = 1u;
                                al += a_step; bh += b_step; bl += b_step;
                            }
                        } else if (!TA && p.a_ring) {
                            // A through tensor memory: one tcgen05.cp per k-step and term moves the 128 x 8 slice of this
                            // CTA's rows (both CTAs of the pair execute it) into 8 columns; the six MMAs of the k-step then
                            // read A there instead of re-reading 6 x 4 KB of shared memory.  cp and mma execute in issue
                            // order (tools/probes/utccp_probe.cu: a ring of two slots reused every other k-step is exact).
                            for (int j = 0; j < ksteps; ++j, ++kstep_n) {
                                const uint32_t t_hi = tmem_base + (uint32_t)p.a_ring + 16u * (kstep_n % 3u), t_lo = t_hi + 8u;
                                utccp2_128x256b(t_hi, ah, a_dhi);
                                utccp2_128x256b(t_lo, al, a_dhi);
                                umma2_tf32_ts(tacc, t_lo, bh, b_dhi, idesc0, acc);
                                umma2_tf32_ts(tacc, t_hi, bl, b_dhi, idesc0, 1u);
                                umma2_tf32_ts(tacc, t_hi, bh, b_dhi, idesc0, 1u);
                                if (p.nr1 > 0) {
                                    umma2_tf32_ts(tacc + p.nr0, t_lo, bh + n1, b_dhi, idesc1, acc);
                                    umma2_tf32_ts(tacc + p.nr0, t_hi, bl + n1, b_dhi, idesc1, 1u);
                                    umma2_tf32_ts(tacc + p.nr0, t_hi, bh + n1, b_dhi, idesc1, 1u);
                                }
                                acc = 1u;
                                ah += a_step; al += a_step; bh += b_step; bl += b_step;
                            }
                        } else
                        for (int j = 0; j < ksteps; ++j) {
                            umma2_tf32(tacc, al, a_dhi, bh, b_dhi, idesc0, acc);                  // small terms first
                            umma2_tf32(tacc, ah, a_dhi, bl, b_dhi, idesc0, 1u);
                            umma2_tf32(tacc, ah, a_dhi, bh, b_dhi, idesc0, 1u);
                            if (p.nr1 > 0) {
                                umma2_tf32(tacc + p.nr0, al, a_dhi, bh + n1, b_dhi, idesc1, acc);
                                umma2_tf32(tacc + p.nr0, ah, a_dhi, bl + n1, b_dhi, idesc1, 1u);
                                umma2_tf32(tacc + p.nr0, ah, a_dhi, bh + n1, b_dhi, idesc1, 1u);
                            }
                            acc = 1u;
                            ah += a_step; al += a_step; bh += b_step; bl += b_step;
                        }
                        umma_commit2(bar_raw_empty + 8 * rs);                // both CTAs: raw slot free once these MMAs retire
                        umma_commit2(bar_lo_empty + 8 * ls);                 //            lo slot free
                        T3_STAMP(1, it, 2);
                    }
                }
                umma_commit2(bar_tmem_full);                                 // accumulator of this tile complete
            }
        }
        __syncwarp();
    } else if (warp < 2 + CONV_WARPS) {
        // ===== converters: lo = tf32(x - trunc(x)), shared memory -> shared memory at the same offset
        const int ct = tid - 64;                                             // 0 .. 255
        const uint32_t conv_leader = map_to_cta(bar_conv_full, 0);
        const int b_vec = p.b_bytes >> 4;                                    // float4s of this CTA's B tile
        uint32_t it = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int z = tile / tiles_mn;
            const int nkbt = (REDUCE ? (p.nsub - z + p.splits - 1) / p.splits : 1) * p.nkb;
            for (int kbt = 0; kbt < nkbt; ++kbt, ++it) {
                const uint32_t rs = it % RAW_STAGES, ruse = it / RAW_STAGES, ls = it % LO_STAGES, luse = it / LO_STAGES;
                const float4* src = reinterpret_cast<const float4*>(raw + rs * SLOT_BYTES);
                float4* dst = reinterpret_cast<float4*>(lo + ls * SLOT_BYTES);
                if (ct == 0) T3_STAMP(2, it, 0);
                mbar_wait(bar_raw_full + 8 * rs, ruse & 1);                  // TMA bytes of this CTA landed
                if (ct == 0) T3_STAMP(2, it, 1);
                auto lo_of = [](float x) { return to_tf32(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); };
                if (p.a_hi_col) {
                    // A: thread = (row 32 q + lane, k half h); the raw values go to tensor memory as the hi term (the MMA
                    // ignores their low mantissa bits), lo back to shared memory at the same place.  B: linear sweep.
                    const int q = warp & 3, h = (warp - 2) >> 2, row = 32 * q + lane;
                    uint32_t a[16];
                    float4 v[5];
                    const uint8_t* araw = raw + rs * SLOT_BYTES;
                    uint8_t* alo = lo + ls * SLOT_BYTES;
                    if (!TA) {                                               // SWIZZLE_128B K-major: 16-byte chunk c of row r at (c ^ (r & 7))
                        const uint32_t rbase = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint4 t = *reinterpret_cast<const uint4*>(araw + rbase + (((4 * h + c) ^ (row & 7)) << 4));
                            a[4 * c] = t.x; a[4 * c + 1] = t.y; a[4 * c + 2] = t.z; a[4 * c + 3] = t.w;
                        }
                    } else {                                                 // 128B_ATOM_32B MN-major: block q, k-row kk, 32-byte chunk (lane >> 3) ^ (kk & 3)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int kk = 16 * h + i;
                            a[i] = *reinterpret_cast<const uint32_t*>(araw + q * 4096 + kk * 128 + ((((lane >> 3) ^ (kk & 3))) << 5) + ((lane & 7) << 2));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const int e = ct + 256 * i;
                        v[i] = e < b_vec ? src[A_TILE / 16 + e] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (luse > 0) mbar_wait(bar_lo_empty + 8 * ls, (luse - 1) & 1);                 // MMA drained the lo slot and the A-hi stage
                    if (ct == 0) T3_STAMP(2, it, 2);
                    tc_fence_after();
                    tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_hi_col + 32u * ls + 16u * h, a);
                    if (!TA) {
                        const uint32_t rbase = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<float4*>(alo + rbase + (((4 * h + c) ^ (row & 7)) << 4)) =
                                make_float4(lo_of(__uint_as_float(a[4 * c])), lo_of(__uint_as_float(a[4 * c + 1])),
                                            lo_of(__uint_as_float(a[4 * c + 2])), lo_of(__uint_as_float(a[4 * c + 3])));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int kk = 16 * h + i;
                            *reinterpret_cast<float*>(alo + q * 4096 + kk * 128 + ((((lane >> 3) ^ (kk & 3))) << 5) + ((lane & 7) << 2)) = lo_of(__uint_as_float(a[i]));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const int e = ct + 256 * i;
                        if (e < b_vec) dst[A_TILE / 16 + e] = make_float4(lo_of(v[i].x), lo_of(v[i].y), lo_of(v[i].z), lo_of(v[i].w));
                    }
                    tmem_st_wait();
                    tc_fence_before();
                } else {
                float4 v[9];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = src[ct + 256 * i];                               // A: 1024 float4
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int e = ct + 256 * i;
                    v[4 + i] = e < b_vec ? src[A_TILE / 16 + e] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (luse > 0) mbar_wait(bar_lo_empty + 8 * ls, (luse - 1) & 1);                     // MMA drained the lo slot
                if (ct == 0) T3_STAMP(2, it, 2);
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    float4 x = v[i], l;
                    l.x = to_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
                    l.y = to_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
                    l.z = to_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
                    l.w = to_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                    if (i < 4) dst[ct + 256 * i] = l;
                    else { const int e = ct + 256 * (i - 4); if (e < b_vec) dst[A_TILE / 16 + e] = l; }
                }
                }
                if (ct == 0) T3_STAMP(2, it, 4);
                fence_async_smem();                                          // generic-proxy writes -> async proxy (UMMA)
                if (ct == 0) T3_STAMP(2, it, 5);
                __syncwarp();
                if (lane == 0) { if (p.cluster_release) mbar_arrive_cluster(conv_leader + 8 * ls); else mbar_arrive_remote(conv_leader + 8 * ls); }    // the leader's conv_full counts the warps of both CTAs
                if (ct == 0) T3_STAMP(2, it, 3);
            }
        }
    } else {
        // ===== epilogue: tensor memory -> registers -> swizzled staging block -> TMA store
        const int q = warp & 3;                                              // tensor-memory lane quarter of this warp
        uint8_t* my_stg = stg + (warp - 2 - CONV_WARPS) * 2 * STG_BYTES;
        const uint32_t empty_leader = map_to_cta(bar_tmem_empty, 0);
        const int chunks_full = (min(p.N, p.n_tile) + 31) >> 5;             // chunks of a full column tile
        const int npad = p.nr0 + p.nr1;
        // chunks of an even tile (accumulator at column 0) that lie in the columns the next (odd) tile accumulates
        // into are drained first: [acc_base1 / 32, chunks); an odd tile shares its first (npad - acc_base1) columns
        const int even_first_full = min(chunks_full, p.acc_base1 >> 5);
        const int odd_cnt_full = min(chunks_full, (max(npad - p.acc_base1, 0) + 31) >> 5);
        uint32_t tile_n = 0, nstore = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tile_n) {
            const int z = tile / tiles_mn, mn = tile - z * tiles_mn, mp = mn / p.tiles_n;
            const int ncol0 = (mn - mp * p.tiles_n) * p.n_tile;              // first output column of this tile
            const int chunks = min(chunks_full, (p.N - ncol0 + 31) >> 5);    // the last column tile may be ragged
            const int even_first = min(chunks, even_first_full), even_cnt = chunks - even_first, odd_cnt = min(chunks, odd_cnt_full);
            const int row0 = mp * 256 + (int)rank * BM + q * 32;             // first output row of this warp
            const float alpha = REDUCE ? 1.0f : p.alpha;
            const int epi = REDUCE ? GSR_EPI_NONE : p.epilogue;
            // The addend is read COALESCED, not by accumulator row: lane -> (row 4 i + lane / 8, 16-byte chunk lane % 8) of the
            // 32 x 32 block, eight lanes per 128-byte row segment, and added in a second pass over the staging block.  (One
            // row per lane -- the accumulator's own layout -- is 32 different lines per load and made the two products of the
            // backward pass that carry an addend three times as slow as the others: 676 against 210-220 us.)  Its blocks are
            // pulled into L2 a tile ahead by cp.async.bulk.prefetch.tensor, one block per lane.
            if (epi == GSR_EPI_ADD) {
                const int nxt = tile + n_clusters;
                if (tile_n == 0 && lane < chunks) tma_prefetch(&mapE, ncol0 + 32 * lane, row0, p.s_extra ? z : 0);
                if (nxt < total_tiles) {
                    const int zn = nxt / tiles_mn, mnn = nxt - zn * tiles_mn, mpn = mnn / p.tiles_n;
                    const int nc0 = (mnn - mpn * p.tiles_n) * p.n_tile;
                    if (lane < min(chunks_full, (p.N - nc0 + 31) >> 5))
                        tma_prefetch(&mapE, nc0 + 32 * lane, mpn * 256 + (int)rank * BM + q * 32, p.s_extra ? zn : 0);
                }
            }
            const bool odd = tile_n & 1;
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (odd ? (uint32_t)p.acc_base1 : 0u);
            const int first_cnt = odd ? odd_cnt : even_cnt;
            auto chunk_at = [&](int i) { return odd ? i : (i < even_cnt ? even_first + i : i - even_cnt); };
            auto release = [&]() {                                           // this warp is done with the shared columns
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (p.cluster_release) mbar_arrive_cluster(empty_leader); else mbar_arrive_remote(empty_leader); }
                if (q == 0 && lane == 0) T3_STAMP(3, tile_n, 2);
            };
            auto emit = [&](uint32_t (&v)[32], int c) {                      // alpha / bias / addend, staging block, TMA store
                const int n0 = ncol0 + c * 32;
                uint8_t* blk = my_stg + (nstore & 1) * STG_BYTES;
                if (nstore >= 2) { if (lane == 0) tma_wait_read<1>(); __syncwarp(); }     // the store that last read this block is done
#pragma unroll
                for (int j = 0; j < 8; ++j) {                                // row = lane; 16-byte chunk j at (j ^ (lane & 7)): conflict free
                    float4 f = make_float4(__uint_as_float(v[4 * j]) * alpha, __uint_as_float(v[4 * j + 1]) * alpha,
                                           __uint_as_float(v[4 * j + 2]) * alpha, __uint_as_float(v[4 * j + 3]) * alpha);
                    if (epi == GSR_EPI_BIAS && n0 + 4 * j < p.N) {           // one address for the whole warp
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.extra + n0 + 4 * j));
                        f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
                    }
                    *reinterpret_cast<float4*>(blk + lane * 128 + ((j ^ (lane & 7)) << 4)) = f;
                }
                if (epi == GSR_EPI_ADD) {
                    // loaded AFTER the staging pass: the accumulator registers are dead by now (loading earlier spills), and the
                    // prefetch has made these L2 hits
                    const int er = lane >> 3, ec = lane & 7;
                    const float* extra_blk = p.extra + (size_t)z * p.s_extra + (size_t)(row0 + er) * p.ld_extra + 4 * ec;
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                            // two half passes of 16 rows: 16 registers of addend
                        float4 e4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            e4[i] = (row0 + 16 * h + 4 * i + er < p.M && n0 + 4 * ec < p.N)          // N % 4 == 0 (launcher)
                                        ? __ldg(reinterpret_cast<const float4*>(extra_blk + (size_t)(16 * h + 4 * i) * p.ld_extra + n0))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {                        // row r, chunk ec: a quarter warp covers one 128-byte row
                            const int r = 16 * h + 4 * i + er;
                            float4* pos = reinterpret_cast<float4*>(blk + r * 128 + ((ec ^ (r & 7)) << 4));
                            float4 f = *pos;
                            f.x += e4[i].x; f.y += e4[i].y; f.z += e4[i].z; f.w += e4[i].w;
                            *pos = f;
                        }
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { tma_store(&mapC, smem_u32(blk), n0, row0, z); tma_commit(); }
                ++nstore;
            };
            if (q == 0 && lane == 0) T3_STAMP(3, tile_n, 0);
            mbar_wait(bar_tmem_full, tile_n & 1);
            tc_fence_after();
            if (q == 0 && lane == 0) T3_STAMP(3, tile_n, 1);
            if (first_cnt == 0) release();
            // two register sets: the tensor-memory load of chunk i + 1 is in flight while chunk i is staged and stored;
            // the first two chunks (drain order) are both read before anything is staged, so that the accumulator is
            // released a load latency after it completes when the shared columns are <= 2 chunks (K-major B)
            uint32_t va[32], vb[32];
            tmem_ld32_async(tacc + (uint32_t)(chunk_at(0) * 32), va);
            if (chunks > 1) tmem_ld32_async(tacc + (uint32_t)(chunk_at(1) * 32), vb);
            tmem_ld_wait(va);
            if (chunks > 1) tmem_ld_wait(vb);
            if (first_cnt == 1 || first_cnt == 2) release();
            for (int i = 0; i < chunks; i += 2) {
                // here: va = chunk i, vb = chunk i + 1 (if it exists), both complete
                emit(va, chunk_at(i));
                if (q == 0 && lane == 0 && i < 9) T3_STAMP(3, tile_n, 3 + i);
                if (i + 2 < chunks) tmem_ld32_async(tacc + (uint32_t)(chunk_at(i + 2) * 32), va);
                if (i + 1 < chunks) {
                    emit(vb, chunk_at(i + 1));
                    if (q == 0 && lane == 0 && i + 1 < 9) T3_STAMP(3, tile_n, 3 + i + 1);
                }
                if (i + 3 < chunks) tmem_ld32_async(tacc + (uint32_t)(chunk_at(i + 3) * 32), vb);
                if (i + 2 < chunks) {
                    tmem_ld_wait(va);
                    if (i + 3 == first_cnt) release();
                    if (i + 3 < chunks) tmem_ld_wait(vb);
                    if (i + 4 == first_cnt) release();
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                      // the peer may still read this CTA's operands / signal its barriers
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = []() -> EncodeFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}

// {cols, rows, items} view of a row-major fp32 operand; strides in elements.
static bool make_map(CUtensorMap* m, const float* base, long long cols, long long rows, long long items, long long ld,
                     long long stride_item, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
    EncodeFn enc = encode_fn();
    if (!enc) return false;
    if (items < 1) items = 1;
    if (stride_item <= 0) stride_item = rows * ld;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)items};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)stride_item * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Measurement aid (gsr_tf32_probe): every CTA pair issues `iters` x 4 tcgen05.mma.cta_group::2 kind::tf32 of
// M 256 x N 256 x K 8 on one resident (zeroed) stage -- the issue rate of the tensor pipe with no operand traffic.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tf32_probe_kernel(int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    for (int e = tid; e < 2 * 128 * 32; e += 128) reinterpret_cast<float*>(smem)[e] = 0.f;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    fence_async_smem();
    if (warp == 0) tmem_alloc2(smem_u32(&slot), 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = instr_desc2(256, 0, 0), dhi = desc_hi(1024, 2);
        const uint32_t a0 = desc_lo(smem_u32(smem), 16), b0 = desc_lo(smem_u32(smem + 128 * 32 * 4), 16);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma2_tf32(tmem, a0 + 2 * j, dhi, b0 + 2 * j, dhi, idesc, 1u);
        }
        umma_commit2(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

}  // namespace t3

bool tma_ok(const Params& p, int transA, int transB) {
    (void)transA;
    if (!g_tma || !t3::encode_fn()) return false;
    if (tma_head_tiles(p.M) < 1 || p.K < 1 || p.N % 4 != 0) return false;
    // a persistent CTA-pair kernel is the wrong tool for a handful of tiles (the Graph U-Net's weight gradients: ONE
    // 268 x k x 268 product, 15 us on three one-CTA tiles against 40 us + a ragged-row launch here)
    if ((long long)p.batch * tma_head_tiles(p.M) * ((p.N + 255) / 256) < 4) return false;
    if (p.lda % 4 || p.ldb % 4 || p.ldc % 4 || p.sA % 4 || p.sB % 4 || p.sC % 4) return false;
    if (!aligned16(p.A) || !aligned16(p.B) || !aligned16(p.C)) return false;
    if (p.epilogue != GSR_EPI_NONE && !aligned16(p.extra)) return false;
    if (p.epilogue == GSR_EPI_ADD && (p.ld_extra % 4 || p.s_extra % 4)) return false;
    return true;
}

int tma_launch(const Params& p, int transA, int transB, bool reduce, int zdim, cudaStream_t st, const Params* second) {
    using namespace t3;
    Plan pl{};
    pl.tiles_m = tma_head_tiles(p.M);
    pl.n_tile = p.N <= 272 ? 272 : 256;
    pl.tiles_n = p.N <= 272 ? 1 : (p.N + 255) / 256;
    pl.items = zdim;
    pl.nsub = reduce ? p.batch : 1;
    pl.splits = reduce ? p.splits : 1;
    pl.M = p.M; pl.N = p.N; pl.K = p.K; pl.nkb1 = (p.K + BK - 1) / BK;
    pl.K2 = second ? second->K : 0;
    pl.nkb = pl.nkb1 + (pl.K2 + BK - 1) / BK;
    // Accumulator column ranges: UMMA N <= 256 per instruction, N / 2 per CTA a multiple of 8.  K-major B: 160 + 112 for
    // the 268-wide products (the split round 1 probed).  MN-major B lives in 32-column blocks (SWIZZLE_128B_BASE32B
    // atoms); a range may end inside a block -- 256 + 16 for 268 columns, the second range reading the first 8 columns
    // of one block per CTA -- unless gsr_set_tuning("tcgemm_mn_partial", 0) restores whole blocks (192 + 128).
    const int npad16 = (std::min(p.N, pl.n_tile) + 15) & ~15;
    if (transB) {
        if (npad16 <= 256) { pl.nr0 = npad16; pl.nr1 = 0; } else { pl.nr0 = 160; pl.nr1 = npad16 - 160; }
    } else if (g_mn_partial) {
        if (npad16 <= 256) { pl.nr0 = npad16; pl.nr1 = 0; } else { pl.nr0 = 256; pl.nr1 = npad16 - 256; }
    } else {
        const int npad = (std::min(p.N, pl.n_tile) + 63) & ~63;
        if (npad <= 256) { pl.nr0 = npad; pl.nr1 = 0; } else { pl.nr0 = 192; pl.nr1 = npad - 192; }
    }
    {
        const int r0 = pl.nr0 / 2, r1 = pl.nr1 / 2;
        if (transB) { pl.b_off1 = r0 * 128; pl.b_bytes = (r0 + r1) * 128; }
        else { pl.b_off1 = ((r0 + 31) / 32) * 4096; pl.b_bytes = pl.b_off1 + ((r1 + 31) / 32) * 4096; }
    }
    // Tensor-memory plan: even tiles accumulate at column 0, odd tiles at acc_base1, the A ring (3 k-steps x (8 hi + 8 lo)
    // columns) behind them: 272 columns -> [0, 272) / [192, 464) / ring at 464; <= 232 columns -> no overlap at all
    // (tcgemm_tmem_a: two 32-column A-hi stages at [448, 512) instead: 272 columns -> [0, 272) / [160, 432))
    pl.a_hi_col = g_tmem_a ? TMEM_COLS - 64 : 0;
    pl.a_ring = (!pl.a_hi_col && !transA && g_utccp) ? TMEM_COLS - 48 : 0;
    pl.acc_base1 = std::min(256, ((pl.a_hi_col ? pl.a_hi_col : pl.a_ring ? pl.a_ring : TMEM_COLS) - (pl.nr0 + pl.nr1)) & ~31);
    pl.pf_ahead = g_pf_ahead;
    pl.cluster_release = g_cluster_release;
    pl.a_batched = p.sA != 0; pl.b_batched = p.sB != 0;
    pl.alpha = p.alpha; pl.epilogue = p.epilogue; pl.extra = p.extra; pl.ld_extra = p.ld_extra; pl.s_extra = p.s_extra;
    const int R0 = pl.nr0 / 2, R1 = pl.nr1 / 2;
    const long long a_items = pl.a_batched ? p.batch : 1, b_items = pl.b_batched ? p.batch : 1;
    CUtensorMap mA, mB0, mB1, mC, mE;
    bool ok = true;
    // op(A) rows = M; A stored [M][K] (K-major) or [K][M] (MN-major)
    if (!transA) ok = ok && make_map(&mA, p.A, p.K, p.M, a_items, p.lda, p.sA, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    else         ok = ok && make_map(&mA, p.A, p.M, p.K, a_items, p.lda, p.sA, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (transB) {                                   // B stored [N][K]
        ok = ok && make_map(&mB0, p.B, p.K, p.N, b_items, p.ldb, p.sB, 32, R0, CU_TENSOR_MAP_SWIZZLE_128B);
        ok = ok && make_map(&mB1, p.B, p.K, p.N, b_items, p.ldb, p.sB, 32, R1 > 0 ? R1 : 8, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {                                        // B stored [K][N]
        ok = ok && make_map(&mB0, p.B, p.N, p.K, b_items, p.ldb, p.sB, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        mB1 = mB0;
    }
    CUtensorMap mA2 = mA, mB02 = mB0, mB12 = mB1;
    if (second) {                                   // the second pair of operands: same shapes and majors, its own K, pitches and bases
        const Params& s = *second;
        if (!transA) ok = ok && make_map(&mA2, s.A, s.K, p.M, a_items, s.lda, s.sA, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        else         ok = ok && make_map(&mA2, s.A, p.M, s.K, a_items, s.lda, s.sA, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (transB) {
            ok = ok && make_map(&mB02, s.B, s.K, p.N, b_items, s.ldb, s.sB, 32, R0, CU_TENSOR_MAP_SWIZZLE_128B);
            ok = ok && make_map(&mB12, s.B, s.K, p.N, b_items, s.ldb, s.sB, 32, R1 > 0 ? R1 : 8, CU_TENSOR_MAP_SWIZZLE_128B);
        } else {
            ok = ok && make_map(&mB02, s.B, p.N, s.K, b_items, s.ldb, s.sB, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
            mB12 = mB02;
        }
    }
    // C: [items][M][N]; REDUCE writes partial sums of pitch sC = M * N, one per split
    ok = ok && make_map(&mC, p.C, p.N, p.M, zdim, reduce ? p.N : p.ldc, p.sC, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    // the addend of GSR_EPI_ADD, seen by the epilogue's L2 prefetch only (32 x 32 blocks, like the stores)
    if (!reduce && p.epilogue == GSR_EPI_ADD)
        ok = ok && make_map(&mE, p.extra, p.N, p.M, p.s_extra ? zdim : 1, p.ld_extra, p.s_extra, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    else mE = mC;
    if (!ok) return fail(GSR_E_UNSUPPORTED, "tcgemm (TMA): cuTensorMapEncodeTiled failed");
    const int total_tiles = pl.items * pl.tiles_m * pl.tiles_n;
    const int clusters = std::min(total_tiles, kNumSMs / 2);
    dim3 grid(2 * clusters, 1, 1);
#define GSR_T3_LAUNCH(TA, TB, RED)                                                                                \
    do {                                                                                                          \
        auto k = tcgemm3_kernel<TA, TB, RED>;                                                                     \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);         \
        if (e != cudaSuccess) return fail((int)e, "tcgemm (TMA): cudaFuncSetAttribute failed");                  \
        k<<<grid, THREADS, SMEM_BYTES, st>>>(mA, mB0, mB1, mC, mE, mA2, mB02, mB12, pl);                                               \
    } while (0)
#define GSR_T3_TT(RED)                                                                                            \
    do {                                                                                                          \
        if (!transA && !transB) GSR_T3_LAUNCH(false, false, RED);                                                 \
        else if (!transA && transB) GSR_T3_LAUNCH(false, true, RED);                                              \
        else if (transA && !transB) GSR_T3_LAUNCH(true, false, RED);                                              \
        else GSR_T3_LAUNCH(true, true, RED);                                                                      \
    } while (0)
    if (reduce) GSR_T3_TT(true); else GSR_T3_TT(false);
#undef GSR_T3_TT
#undef GSR_T3_LAUNCH
    return launched(reduce ? "tcgemm3_reduce" : "tcgemm3");
}

}  // namespace tc
}  // namespace gsr

#ifdef GSR_T3_TRACE
extern "C" __attribute__((visibility("default"))) int gsr_tcgemm3_trace_read(unsigned long long* host, int count) {
    return (int)cudaMemcpyFromSymbol(host, gsr::tc::t3::g_trace, sizeof(unsigned long long) * count);
}
#endif

extern "C" double gsr_tf32_probe(int iters, void* stream) {
    using namespace gsr;
    if (iters <= 0) return 0.0;
    const int smem = 2 * 128 * 32 * 4 + 1024;
    if (cudaFuncSetAttribute(tc::t3::tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 0.0;
    tc::t3::tf32_probe_kernel<<<kNumSMs, 128, smem, (cudaStream_t)stream>>>(iters);
    if (launched("tf32_probe")) return 0.0;
    return (double)(kNumSMs / 2) * iters * 4.0 * 2.0 * 256.0 * 256.0 * 8.0;
}
