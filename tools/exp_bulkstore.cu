// Microbenchmark: how fast can CTAs stream zero tiles to HBM (a) with cp.async.bulk shared->global, (b) with st.global.v4?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
template<int RING>
__global__ void bulk_kernel(float* out, long long ntiles, int tile_floats, int sync_mode, int order=0)
{
    extern __shared__ __align__(128) float s[];
    for (int i = threadIdx.x; i < RING*tile_floats; i += blockDim.x) s[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    int slot = 0;
    long long lo = 0, hi = ntiles, step = 1;
    if (order == 0) { lo = blockIdx.x; step = gridDim.x; }
    else { lo = ntiles * blockIdx.x / gridDim.x; hi = ntiles * (blockIdx.x + 1) / gridDim.x; }
    for (long long tt = lo; tt < hi; tt += step) {
        long long t = tt;
        if (order == 2) { long long plane = tt % 72, bt = tt / 72; t = (bt / 4) * 288 + plane * 4 + (bt % 4); if (t >= ntiles) t = tt; }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(RING-1) : "memory");
        if (sync_mode) { __syncthreads(); s[slot*tile_floats + threadIdx.x] = (float)t; asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); __syncthreads(); }
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + t*tile_floats), "r"(smem_u32(s + slot*tile_floats)), "r"(tile_floats*4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        slot = (slot+1==RING)?0:slot+1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void st_kernel(float4* out, long long n4)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) out[i] = make_float4(0,0,0,0);
}
int main(){
    const long long bytes = 281LL*1024*1024;
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
        printf("%-50s %.1f us  %.0f GB/s  (%s)\n", name, ms/20*1e3, bytes/(ms/20*1e-3)/1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int tile_kb : {8, 16, 32, 64}) for (int cps : {1, 2, 4}) for (int sync_mode : {0,1}) {
        int tf = tile_kb*256; long long nt = bytes/4/tf; size_t sm = 3*tf*4;
        if (sm*cps > 220*1024) continue;
        cudaFuncSetAttribute(bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        char nm[128]; snprintf(nm,128,"bulk ring3 tile=%dKB ctas/sm=%d sync=%d", tile_kb, cps, sync_mode);
        run(nm, [&](float* b){ bulk_kernel<3><<<sms*cps,256,sm>>>(b, nt, tf, sync_mode); });
    }
    for (int tile_kb : {16, 32}) for (int cps : {1, 2}) {
        int tf = tile_kb*256; long long nt = bytes/4/tf; size_t sm = 6*tf*4;
        if (sm*cps > 220*1024) continue;
        cudaFuncSetAttribute(bulk_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        char nm[128]; snprintf(nm,128,"bulk ring6 tile=%dKB ctas/sm=%d sync=1", tile_kb, cps);
        run(nm, [&](float* b){ bulk_kernel<6><<<sms*cps,256,sm>>>(b, nt, tf, 1); });
    }
    for (int order : {0, 1, 2}) {
        int tf = 32*256; long long nt = bytes/4/tf; size_t sm = 3*tf*4;
        cudaFuncSetAttribute(bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        char nm[128]; snprintf(nm,128,"bulk ring3 tile=32KB ctas/sm=2 sync=1 order=%d", order);
        run(nm, [&](float* b){ bulk_kernel<3><<<sms*2,256,sm>>>(b, nt, tf, 1, order); });
    }
    for (int mult : {4, 8, 16, 32}) { char nm[128]; snprintf(nm,128,"st.global.v4 grid=%dxSMs x256", mult);
        run(nm, [&](float* b){ st_kernel<<<sms*mult,256>>>((float4*)b, bytes/16); }); }
    return 0;
}
