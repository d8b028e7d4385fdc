// Library-level entry points of libgsr_sm100.so: ABI version, error text, launch counter and
// the flat gradient bucket pack / unpack used by the subject-data-parallel all-reduce.
#include "common.cuh"

namespace gsr {
thread_local char g_last_error[256] = "";
std::atomic<uint64_t> g_launches{0};

// One CTA row per tensor (blockIdx.y), grid-stride over its elements.
template <bool PACK>
__global__ void flat_copy_kernel(float* const* __restrict__ ptrs, const long long* __restrict__ offsets,
                                 float* __restrict__ flat, float scale) {
    const int t = blockIdx.y;
    const long long lo = offsets[t], hi = offsets[t + 1];
    float* p = ptrs[t];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < hi - lo;
         e += (long long)gridDim.x * blockDim.x) {
        if (PACK) flat[lo + e] = p[e] * scale;
        else p[e] = flat[lo + e] * scale;
    }
}
}  // namespace gsr

using namespace gsr;

extern "C" int gsr_abi_version(void) { return GSR_ABI_VERSION; }
extern "C" const char* gsr_last_error(void) { return g_last_error; }
extern "C" uint64_t gsr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int gsr_flat_pack(const float* const* ptrs_dev, const long long* offsets_dev, int n_tensors,
                             float* flat, long long total, float scale, void* stream) {
    if (!ptrs_dev || !offsets_dev || !flat || n_tensors <= 0 || total <= 0) return fail(GSR_E_ARG, "flat_pack: bad argument");
    flat_copy_kernel<true><<<dim3(32, n_tensors), 256, 0, (cudaStream_t)stream>>>(
        const_cast<float* const*>(reinterpret_cast<const float* const*>(ptrs_dev)), offsets_dev, flat, scale);
    return launched("flat_pack");
}

extern "C" int gsr_flat_unpack(float* const* ptrs_dev, const long long* offsets_dev, int n_tensors,
                               const float* flat, long long total, float scale, void* stream) {
    if (!ptrs_dev || !offsets_dev || !flat || n_tensors <= 0 || total <= 0) return fail(GSR_E_ARG, "flat_unpack: bad argument");
    flat_copy_kernel<false><<<dim3(32, n_tensors), 256, 0, (cudaStream_t)stream>>>(
        ptrs_dev, offsets_dev, const_cast<float*>(flat), scale);
    return launched("flat_unpack");
}
