// Microbenchmark for the grid writer: warps that each run their OWN ring of shared-memory tiles and bulk stores
// (no CTA barrier anywhere) against the CTA-synchronous form.  Streams 302 MB of mostly-zero tiles to HBM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_fillwarp tools/exp_fillwarp.cu && tools/exp_fillwarp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
template<int RW>
__global__ void warp_kernel(float* out, long long ntiles, int tile_floats, int cells_per_lane)
{
    extern __shared__ __align__(128) float s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* ring = s + (size_t)wid * RW * tile_floats;
    for (int i = lane; i < RW * tile_floats; i += 32) ring[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    int slot = 0;
    const long long gw = (long long)blockIdx.x * nw + wid, tw = (long long)gridDim.x * nw;
    for (long long t = gw; t < ntiles; t += tw) {
        float* tile = ring + slot * tile_floats;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(RW-1) : "memory");
        __syncwarp();
        for (int c = 0; c < cells_per_lane; ++c) tile[(lane * 37 + c * 1031) % tile_floats] = (float)t;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + t*tile_floats), "r"(smem_u32(tile)), "r"(tile_floats*4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        slot = (slot+1==RW)?0:slot+1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main(){
    const long long bytes = 288LL*1024*1024;
    float* buf[3]; for (int i=0;i<3;i++) cudaMalloc(&buf[i], bytes);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto launch){
        for (int i=0;i<3;i++) launch(buf[i%3]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i=0;i<20;i++) launch(buf[i%3]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms,e0,e1);
        printf("%-60s %.1f us  %.0f GB/s  (%s)\n", name, ms/20*1e3, bytes/(ms/20*1e-3)/1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int tile_kb : {4, 8, 16, 32}) for (int warps : {2, 4, 6, 8, 12}) for (int rw : {2, 3}) {
        int tf = tile_kb*256; long long nt = bytes/4/tf; size_t sm = (size_t)warps*rw*tf*4;
        if (sm > 200*1024) continue;
        char nm[128]; snprintf(nm,128,"warp rings: tile=%dKB warps=%d slots/warp=%d smem=%zuKB", tile_kb, warps, rw, sm/1024);
        if (rw == 2) { cudaFuncSetAttribute(warp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            run(nm, [&](float* b){ warp_kernel<2><<<sms,warps*32,sm>>>(b, nt, tf, 4); }); }
        else { cudaFuncSetAttribute(warp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            run(nm, [&](float* b){ warp_kernel<3><<<sms,warps*32,sm>>>(b, nt, tf, 4); }); }
    }
    return 0;
}
