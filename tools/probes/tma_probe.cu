// Probe (round 2): what tcgen05.mma kind::tf32 needs from TMA-landed fp32 tiles.  Five questions, one binary:
//   T1  does kind::tf32 truncate or round the 13 low mantissa bits of a 32-bit operand container (both operands)?
//       -> decides whether the raw fp32 tile a TMA load lands IS the `hi` term of the 3xTF32 split.
//   T2  K-major operands loaded by cp.async.bulk.tensor with CU_TENSOR_MAP_SWIZZLE_128B (box 32 floats x rows):
//       matrix descriptor layout type 2 (SWIZZLE_128B), SBO = 1024, k-step = +32 bytes on the start address.
//   T3  MN-major operands loaded with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (box 32 floats (m / n) x 32 k-rows, one box
//       per 32-wide block): layout type 1 (SWIZZLE_128B_BASE32B), LBO = 4096 (box pitch), SBO = 512, k-step = +1024.
//   T4  the mixed cases (NN, TT) with the same descriptors.
//   T5  cp.async.bulk.tensor store of a row-major staging tile with clipping at the tensor edge (ragged 268).
//   T6  tcgen05.mma kind::tf32 issue rate with resident operands (cta_group::2, M 256 x N 256 x K 8), all SMs.
// Every wait is bounded; a protocol error prints a flag instead of hanging.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr & 0x3FFFFu) >> 4) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo, uint32_t lt) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((lt & 7u) << 29); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ T1
// Thread-staged K-major no-swizzle tiles (the layout csrc/tcgemm.cu has used since round 1).  A [128][8], B [16][8].
__global__ void __launch_bounds__(128, 1) probe_trunc(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int* flag) {
    __shared__ __align__(128) uint8_t a_tile[128 * 8 * 4];
    __shared__ __align__(128) uint8_t b_tile[16 * 8 * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t SBO = (8 / 4) * 128;     // two 16-byte k chunks per 8-row core-matrix row group
    for (int e = tid; e < 128 * 8; e += 128) {
        const int r = e / 8, k = e % 8;
        *reinterpret_cast<float*>(a_tile + (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4) = A[e];
    }
    for (int e = tid; e < 16 * 8; e += 128) {
        const int r = e / 8, k = e % 8;
        *reinterpret_cast<float*>(b_tile + (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4) = B[e];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        umma_tf32(tmem, desc_lo(smem_u32(a_tile), 128), desc_hi(SBO, 0), desc_lo(smem_u32(b_tile), 128), desc_hi(SBO, 0), idesc, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    if (!mbar_wait_bounded(smem_u32(&bar), 0)) { if (tid == 0) flag[0] = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int c = 0; c < 16; ++c) C[tid * 16 + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32u) : "memory");
}

// ------------------------------------------------------------------------------------------------ T2 - T5
// C [128][64] = op(A) op(B), K = 64 (two k-blocks of 32), operands loaded by TMA, result written by a TMA store
// into a [M_valid][N_valid] tensor (clipped).  A_MN: A stored [K][128] (MN-major) else [128][K]; B_MN: B stored [K][64].
constexpr int PM = 128, PN = 64, PK = 64, PBK = 32;
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128, 1)
probe_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC, int* flag) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                              // 2 k-blocks x 16 KB
    uint8_t* b_tile = smem + 2 * PM * PBK * 4;           // 2 k-blocks x 8 KB
    uint8_t* c_tile = b_tile + 2 * PN * PBK * 4;         // 128 x 64 floats row-major (32 KB)
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (tid == 0) {
        mbar_expect_tx(bar_tma, 2 * (PM + PN) * PBK * 4);
        for (int kb = 0; kb < 2; ++kb) {
            if (!A_MN) tma_load_3d(smem_u32(a_tile + kb * PM * PBK * 4), &mapA, bar_tma, kb * PBK, 0, 0);          // box {32 k, 128 m}
            else for (int j = 0; j < PM / 32; ++j)                                                                   // box {32 m, 32 k} per 32-row block
                tma_load_3d(smem_u32(a_tile + kb * PM * PBK * 4 + j * 4096), &mapA, bar_tma, j * 32, kb * PBK, 0);
            if (!B_MN) tma_load_3d(smem_u32(b_tile + kb * PN * PBK * 4), &mapB, bar_tma, kb * PBK, 0, 0);          // box {32 k, 64 n}
            else for (int j = 0; j < PN / 32; ++j)
                tma_load_3d(smem_u32(b_tile + kb * PN * PBK * 4 + j * 4096), &mapB, bar_tma, j * 32, kb * PBK, 0);
        }
        if (!mbar_wait_bounded(bar_tma, 0)) flag[0] = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(PN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
        // K-major SWIZZLE_128B: 8-row groups of 1024 bytes (SBO), LBO unused (1); k-step (8 floats) = +32 bytes
        // MN-major SWIZZLE_128B_BASE32B: 32-wide blocks 4096 bytes apart (LBO), 4-k groups 512 bytes apart (SBO); k-step = +1024
        const uint32_t a_dhi = A_MN ? desc_hi(512, 1) : desc_hi(1024, 2), b_dhi = B_MN ? desc_hi(512, 1) : desc_hi(1024, 2);
        const uint32_t a_lbo = A_MN ? 4096 : 16, b_lbo = B_MN ? 4096 : 16;
        const uint32_t a_k = A_MN ? 1024 : 32, b_k = B_MN ? 1024 : 32;
        uint32_t acc = 0;
        for (int kb = 0; kb < 2; ++kb)
            for (int j = 0; j < PBK / 8; ++j) {
                umma_tf32(tmem, desc_lo(smem_u32(a_tile + kb * PM * PBK * 4) + j * a_k, a_lbo), a_dhi,
                          desc_lo(smem_u32(b_tile + kb * PN * PBK * 4) + j * b_k, b_lbo), b_dhi, idesc, acc);
                acc = 1;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_mma) : "memory");
    }
    if (!mbar_wait_bounded(bar_mma, 0)) { if (tid == 0) flag[1] = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < PN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int c = 0; c < 32; ++c) reinterpret_cast<uint32_t*>(c_tile)[tid * PN + c0 + c] = v[c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        tma_store_3d(&mapC, smem_u32(c_tile), 0, 0, 0);             // box {64, 128}: clipped to the tensor extents
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64u) : "memory");
}

// ------------------------------------------------------------------------------------------------ T6
// Resident operands: every CTA pair issues `iters` x 4 MMAs (M 256, N 256, K 8) on one zeroed stage; commit every 4.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_rate(int iters, int* flag) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                      // 128 x 32 floats
    uint8_t* b_tile = smem + 128 * 32 * 4;       // 128 x 32 floats (this CTA's half of N = 256)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int e = tid; e < 2 * 128 * 32; e += 128) reinterpret_cast<float*>(smem)[e] = 0.f;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t dhi = desc_hi(1024, 2);
        const uint32_t a0 = desc_lo(smem_u32(a_tile), 16), b0 = desc_lo(smem_u32(b_tile), 16);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                             "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
                             :: "r"(tmem), "r"(a0 + 2 * j), "r"(dhi), "r"(b0 + 2 * j), "r"(dhi), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     :: "r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    {
        uint32_t ok = 0;
        for (uint32_t spin = 0; !ok && spin < (1u << 28); ++spin)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (!ok && tid == 0) flag[0] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static CUtensorMap make_map(const float* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld, uint64_t stride_b,
                            uint32_t box_c, uint32_t box_r, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    cuuint64_t dims[3] = {cols, rows, batch};
    cuuint64_t strides[2] = {ld * 4, stride_b * 4};
    cuuint32_t box[3] = {box_c, box_r, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (cols %llu rows %llu ld %llu box %u x %u sw %d)\n", (int)r,
                                    (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld, box_c, box_r, (int)sw); exit(2); }
    return m;
}

static float trunc13(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna13(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rne13(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <bool A_MN, bool B_MN>
static void run_tma_case(const char* name) {
    const int LD = 268;                                  // the pitch of the real operands (1072 bytes)
    const int Mv = 124, Nv = 60;                         // valid extents of C: rows 124..127 and columns 60..63 must stay untouched
    std::vector<float> A((size_t)(A_MN ? PK : PM) * LD, 0.f), B((size_t)(B_MN ? PK : PN) * LD, 0.f), C((size_t)PM * LD, -7.f);
    auto a_at = [&](int m, int k) -> float& { return A_MN ? A[(size_t)k * LD + m] : A[(size_t)m * LD + k]; };
    auto b_at = [&](int n, int k) -> float& { return B_MN ? B[(size_t)k * LD + n] : B[(size_t)n * LD + k]; };
    for (int m = 0; m < PM; ++m) for (int k = 0; k < PK; ++k) a_at(m, k) = (float)((m * 7 + k * 3 + 1) % 13 - 6);
    for (int n = 0; n < PN; ++n) for (int k = 0; k < PK; ++k) b_at(n, k) = (float)((n * 5 + k * 11 + 2) % 9 - 4);
    float *dA, *dB, *dC; int* dF;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&dF, 16));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dF, 0, 16));
    CUtensorMap mA = A_MN ? make_map(dA, PM, PK, 1, LD, (uint64_t)PK * LD, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                          : make_map(dA, PK, PM, 1, LD, (uint64_t)PM * LD, 32, PM, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mB = B_MN ? make_map(dB, PN, PK, 1, LD, (uint64_t)PK * LD, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                          : make_map(dB, PK, PN, 1, LD, (uint64_t)PN * LD, 32, PN, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mC = make_map(dC, Nv, Mv, 1, LD, (uint64_t)PM * LD, PN, PM, CU_TENSOR_MAP_SWIZZLE_NONE);
    const int smem = 2 * PM * PBK * 4 + 2 * PN * PBK * 4 + PM * PN * 4 + 1024;
    auto k = probe_tma<A_MN, B_MN>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<1, 128, smem>>>(mA, mB, mC, dF);
    cudaError_t e = cudaDeviceSynchronize();
    int F[4]; CK(cudaMemcpy(F, dF, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0; int touched_outside = 0, untouched_inside = 0;
    for (int m = 0; m < PM; ++m) for (int n = 0; n < LD; ++n) {
        const float got = C[(size_t)m * LD + n];
        if (m < Mv && n < Nv) {
            double ref = 0;
            for (int kk = 0; kk < PK; ++kk) ref += (double)a_at(m, kk) * b_at(n, kk);
            err = fmax(err, fabs(ref - got));
            if (got == -7.f && ref != -7.0) ++untouched_inside;
        } else if (got != -7.f) ++touched_outside;
    }
    printf("%s: sync %s, barrier timeouts tma %d mma %d, max |C - ref| = %.3g, elements outside the clipped box touched: %d, inside untouched: %d\n",
           name, cudaGetErrorString(e), F[0], F[1], err, touched_outside, untouched_inside);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dF);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not found\n"); return 2; }
    g_encode = (EncodeFn)fn;

    {   // ---- T1
        std::vector<float> A(128 * 8), B(16 * 8, 0.f), C(128 * 16);
        uint32_t s = 12345u;
        for (auto& x : A) { s = s * 1664525u + 1013904223u; uint32_t u = 0x3F800000u | (s >> 9); memcpy(&x, &u, 4); if (s & 256) x = -x; }
        for (int k = 0; k < 8; ++k) B[k * 8 + k] = 1.0f;                 // B[n][k] = delta(n, k), n < 8 -> C[m][n] = A as read
        float *dA, *dB, *dC; int* dF;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&dF, 16));
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<float> A2 = A, B2 = B;
            if (pass == 1) {                                             // roles swapped: A one-hot (row m selects k = m % 8), B random
                for (auto& x : A2) x = 0.f;
                for (int m = 0; m < 128; ++m) A2[m * 8 + (m % 8)] = 1.0f;
                for (int i = 0; i < 16 * 8; ++i) B2[i] = A[i];
            }
            CK(cudaMemcpy(dA, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(dF, 0, 16));
            probe_trunc<<<1, 128>>>(dA, dB, dC, dF);
            cudaError_t e = cudaDeviceSynchronize();
            int F[4]; CK(cudaMemcpy(F, dF, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
            int n_trunc = 0, n_rna = 0, n_rne = 0, n_exact = 0, total = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < (pass == 0 ? 8 : 16); ++n) {
                const float src = pass == 0 ? A[m * 8 + n] : A[n * 8 + (m % 8)];
                const float got = C[m * 16 + n];
                ++total;
                n_trunc += got == trunc13(src); n_rna += got == rna13(src); n_rne += got == rne13(src); n_exact += got == src;
            }
            printf("T1 %s operand: sync %s, timeout %d; of %d full-mantissa values the MMA read: truncated %d, rna %d, rne %d, all 23 bits %d\n",
                   pass == 0 ? "A" : "B", cudaGetErrorString(e), F[0], total, n_trunc, n_rna, n_rne, n_exact);
            if (pass == 0) printf("   sample: in %.9g -> out %.9g (trunc %.9g, rna %.9g)\n", A[3], C[3], trunc13(A[3]), rna13(A[3]));
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dF);
    }
    run_tma_case<false, false>("T2 NT (A K-major SW128, B K-major SW128)");
    run_tma_case<true, true>("T3 TN (A MN-major ATOM_32B, B MN-major ATOM_32B)");
    run_tma_case<false, true>("T4 NN (A K-major, B MN-major)");
    run_tma_case<true, false>("T4 TT (A MN-major, B K-major)");
    {   // ---- T6
        int* dF; CK(cudaMalloc(&dF, 16)); CK(cudaMemset(dF, 0, 16));
        const int smem = 2 * 128 * 32 * 4 + 1024;
        CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 3; ++rep) {
            const int iters = rep == 0 ? 256 : 8192;
            CK(cudaEventRecord(e0));
            probe_rate<<<148, 128, smem>>>(iters, dF);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
            int F[4]; CK(cudaMemcpy(F, dF, 16, cudaMemcpyDeviceToHost));
            const double flop = 74.0 * iters * 4.0 * 2.0 * 256 * 256 * 8;
            printf("T6 tf32 issue rate (74 CTA pairs, M256 N256 K8 x %d): sync %s, timeout %d, %.3f ms -> %.1f TFLOP/s dense tf32\n",
                   iters * 4, cudaGetErrorString(e), F[0], ms, flop / (ms * 1e-3) / 1e12);
        }
        cudaFree(dF);
    }
    return 0;
}
