// Probe: tcgen05.mma.cta_group::2 (CTA pair, M = 256) with thread-staged operands in the K-major no-swizzle
// canonical layout used by csrc/tcgemm.cu.  Each CTA stages its own 128 rows of A and HALF of B (N/2 rows);
// the leader issues the MMAs; tcgen05.commit multicasts to a barrier in both CTAs; each CTA drains its own
// 128 accumulator rows.  Prints the max error against A.B^T for both candidate B-half orders.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>

constexpr int BM = 128, N = 32, NH = N / 2, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr & 0x3FFFFu) >> 4) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo, uint32_t lt) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((lt & 7u) << 29); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int* __restrict__ flag) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_tile = smem;                         // 128 x 32 floats = 16 KB, K-major: core (r/8, k/4) at (r/8)*SBO + (k/4)*128 + (r%8)*16
    uint8_t* b_tile = smem + BM * K * 4;            // NH x 32 floats
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + NH * K * 4);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t SBO = (K / 4) * 128;
    // stage A rows [128 rank, 128 rank + 128) and B rows [NH rank, NH rank + NH)
    for (int e = tid; e < BM * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = A[(size_t)(rank * BM + r) * K + k];
        *reinterpret_cast<float*>(a_tile + (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4) = v;
    }
    for (int e = tid; e < NH * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = B[(size_t)(rank * NH + r) * K + k];
        *reinterpret_cast<float*>(b_tile + (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4) = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
    if (rank == 0 && tid == 0) {
        // instruction descriptor: fp32 accumulate, tf32 x tf32, both K-major, N, M = 256
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t dhi = desc_hi(SBO, 0);
        uint32_t a_lo = desc_lo(smem_u32(a_tile), 128), b_lo = desc_lo(smem_u32(b_tile), 128);
        for (int j = 0; j < K / 8; ++j) {
            const uint32_t acc = j ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                         "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
                         :: "r"(tmem), "r"(a_lo), "r"(dhi), "r"(b_lo), "r"(dhi), "r"(idesc), "r"(acc) : "memory");
            a_lo += 256 >> 4; b_lo += 256 >> 4;
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
    // everyone waits on the local barrier (bounded)
    {
        uint32_t ok = 0;
        for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        if (!ok) { if (tid == 0) flag[rank] = 1; }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = rank * BM + warp * 32 + lane;
        for (int c = 0; c < N; ++c) C[(size_t)row * N + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32u) : "memory");
}

int main() {
    std::vector<float> A(256 * K), B(N * K), C(256 * N, -1.f);
    for (int i = 0; i < 256 * K; ++i) A[i] = (float)((i * 7 + 3) % 11 - 5);
    for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6);
    float *dA, *dB, *dC; int* dF;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dF, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, C.size() * 4); cudaMemset(dF, 0, 8);
    const int smem = BM * K * 4 + NH * K * 4 + 64;
    probe<<<2, 128, smem>>>(dA, dB, dC, dF);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch/sync: %s\n", cudaGetErrorString(e));
    int F[2]; cudaMemcpy(F, dF, 8, cudaMemcpyDeviceToHost); cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    printf("barrier timeouts: %d %d\n", F[0], F[1]);
    double err_plain = 0, err_swapped = 0;
    for (int r = 0; r < 256; ++r) for (int c = 0; c < N; ++c) {
        double ref = 0, ref_sw = 0;
        const int cs = (c + NH) % N;
        for (int k = 0; k < K; ++k) { ref += (double)A[r * K + k] * B[c * K + k]; ref_sw += (double)A[r * K + k] * B[cs * K + k]; }
        err_plain = fmax(err_plain, fabs(ref - C[r * N + c])); err_swapped = fmax(err_swapped, fabs(ref_sw - C[r * N + c]));
    }
    printf("max |C - A.B^T|: B halves in rank order %.3g, swapped %.3g\n", err_plain, err_swapped);
    printf("C[0][0..3] = %g %g %g %g   C[128][0..3] = %g %g %g %g\n", C[0], C[1], C[2], C[3], C[128 * N], C[128 * N + 1], C[128 * N + 2], C[128 * N + 3]);
    return 0;
}
