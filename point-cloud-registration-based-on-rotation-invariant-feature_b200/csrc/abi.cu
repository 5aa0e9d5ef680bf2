// abi.cu — ABI bookkeeping for libri_b200.so (include/ri_b200.h).
#include "ri_common.cuh"

extern "C" int ri_abi_version(void) { return 1; }

// Debug aid for tools/timeline.py: one thread stores the GPU's nanosecond timer; enqueued between the kernels of a
// step it yields the start/end of every kernel on every stream (a profiler cannot show concurrent branches).
namespace {
__global__ void stamp_kernel(unsigned long long* slot)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
}  // namespace

extern "C" int ri_debug_stamp(unsigned long long* slot, void* stream)
{
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
