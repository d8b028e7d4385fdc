// Shared declarations of the tcgen05 contraction kernels: csrc/tcgemm.cu (thread-staged operands: one CTA per
// tile, CTA pairs, FP32 ragged-row kernels) and csrc/tcgemm_tma.cu (TMA-fed persistent CTA pairs).
#pragma once
#include "common.cuh"

namespace gsr {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;

struct Params {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    long long sA, sB, sC;
    int batch, splits;
    float alpha;
    int epilogue;
    const float* extra; int ld_extra; long long s_extra;
};

extern int g_pairs, g_tail, g_dual;

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
// D[tmem] (+)= A[smem] . B[smem], kind::tf32, issued by ONE thread.  The 64-bit matrix
// descriptors are passed as (low word, high word): only the low word (start address) changes
// from one MMA to the next, so the issuing thread spends one integer add per operand.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "setp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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

// Shared-memory matrix descriptor (sm_100 version bit set), as two 32-bit words.
//   low  word: start address >> 4 in bits [0,14), leading byte offset >> 4 in bits [16,30)
//   high word: stride byte offset >> 4 in bits [0,14), version = 1 at bit 14, layout type in
//              bits [29,32): 0 = SWIZZLE_NONE ("interleave"), 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) {
    return ((addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29);
}
// Instruction descriptor: fp32 accumulate, tf32 x tf32, M = 128, runtime N, operand majors.
__device__ __forceinline__ uint32_t instr_desc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Round-to-nearest (ties away) to the 10-bit TF32 mantissa with two integer ops: add half an
// ulp of the dropped 13 bits to the magnitude, clear them.  (`cvt.rna.tf32.f32` compiles to the
// same pair plus an FSETP / SEL for Inf / NaN, which the split does not need: Inf stays Inf --
// 0x1000 is masked away again -- and NaN stays NaN.)
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}


// ---------------------------------------------------------------- CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// Remote arrive with the default (release, CTA-scope) semantics.  For a producer whose DATA stays in its own CTA's shared
// memory -- already fenced to the async proxy with fence.proxy.async -- and whose consumer is a tcgen05.mma the leader
// issues afterwards: only the arrival itself crosses the CTA boundary.  The cluster-scope release above costs
// 0.6-1.6 us per arrival inside the k-loop (profiles/r2_tcgemm3_pipeline_trace.txt).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait_cluster(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {     // arrives on `bar` of BOTH CTAs once the MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "setp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t instr_desc2(int n, int a_mn_major, int b_mn_major) {      // M = 256
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// csrc/tcgemm_tma.cu: the TMA-fed persistent CTA-pair kernel.  tma_ok: shapes / alignments it takes (the caller has
// already applied pair_ok); tma_launch covers the first 256 * floor(M / 256) rows of every batch item.
extern int g_tma, g_utccp, g_mn_partial, g_pf_ahead, g_tmem_a, g_cluster_release;       // gsr_set_tuning("tcgemm_tma" | "tcgemm_utccp" | "tcgemm_mn_partial", 0|1)
// 256-row tiles the TMA kernel takes: every full one and a last partial one with at least 64 rows (rows past M are
// zero filled by the loads and clipped by the stores); fewer rows than that go to the FP32 ragged-row kernels.
extern int g_partial_head, g_reduce_simt_tails;     // gsr_set_tuning("tcgemm_partial_head" | "tcgemm_reduce_simt_tails", 0|1)
inline int tma_head_tiles(int M) { return M / 256 + ((g_partial_head && M % 256 >= 64) ? 1 : 0); }
inline int tma_head_rows(int M) { const int r = tma_head_tiles(M) * 256; return r < M ? r : M; }
bool tma_ok(const Params& p, int transA, int transB);
// `second` (A, B, K, lda, ldb, sA, sB of a second pair of operands with the same shapes and majors): C = alpha (op(A) op(B) + op(A2) op(B2)) + ...,
// both products accumulated in tensor memory (the k-loop runs over the first pair's k-blocks, then the second's).
int tma_launch(const Params& p, int transA, int transB, bool reduce, int zdim, cudaStream_t st, const Params* second = nullptr);

}  // namespace tc
}  // namespace gsr
