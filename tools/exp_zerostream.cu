// Microbenchmark: how fast can zeros be streamed to HBM by cp.async.bulk from ONE constant shared-memory tile, and what does
// it cost to patch a few cells of every tile afterwards with ordinary stores?  (Grid writer, zero-stream form: csrc/voxelize.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_zerostream tools/exp_zerostream.cu && tools/exp_zerostream
// Modes
//   0  per-warp ring of RW patched tiles, flow control by wait_group.read (the ring form's pattern)
//   1  constant tile, paced: after every bulk store wait_group.read <= K pending
//   2  constant tile, unpaced: every tile issued back to back, one wait at the end
//   3  constant tile, patch afterwards: full completion wait with D tiles in flight, then `cells` 4-byte stores per lane into the tile
//   4  plain st.global.v4 zero stores (32 lanes x 16 B per instruction)
// Address pattern: 0 = tile t at t * tile (all warps write neighbouring tiles), 1 = the grid's: 8 planes of one tile 128 KB apart
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void wait_full() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ long long tile_addr(long long t, int tile_floats, int pattern, long long plane_floats)
{
    if (pattern == 0) return t * tile_floats;
    // grid pattern: item = 8 planes of one tile; planes are plane_floats apart, tiles of a plane are contiguous
    const long long tiles_per_plane = plane_floats / tile_floats;
    const long long p = t & 7, it = t >> 3;
    const long long tile = it % tiles_per_plane, grp = it / tiles_per_plane;
    return (grp * 8 + p) * plane_floats + tile * tile_floats;
}

template <int MODE, int K>
__global__ void zs_kernel(float* out, long long ntiles, int tile_floats, int cells, int pattern, long long plane_floats)
{
    extern __shared__ __align__(128) float s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * nw + wid, tw = (long long)gridDim.x * nw;
    if constexpr (MODE == 0) {
        float* ring = s + (size_t)wid * K * tile_floats;
        for (int i = lane; i < K * tile_floats; i += 32) ring[i] = 0.f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        int slot = 0;
        for (long long t = gw; t < ntiles; t += tw) {
            float* tile = ring + slot * tile_floats;
            if (lane == 0) wait_read<K - 1>();
            __syncwarp();
            for (int c = 0; c < cells; ++c) tile[(lane * 37 + c * 1031) % tile_floats] = (float)t;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) bulk_store(out + tile_addr(t, tile_floats, pattern, plane_floats), tile, tile_floats * 4);
            slot = (slot + 1 == K) ? 0 : slot + 1;
        }
        if (lane == 0) wait_full<0>();
        return;
    }
    if constexpr (MODE == 4) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long t = gw; t < ntiles; t += tw) {
            float4* dst = reinterpret_cast<float4*>(out + tile_addr(t, tile_floats, pattern, plane_floats));
            for (int i = lane; i < tile_floats / 4; i += 32) dst[i] = z;
        }
        return;
    }
    for (int i = threadIdx.x; i < tile_floats; i += blockDim.x) s[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if constexpr (MODE == 1) {
        for (long long t = gw; t < ntiles; t += tw)
            if (lane == 0) { bulk_store(out + tile_addr(t, tile_floats, pattern, plane_floats), s, tile_floats * 4); wait_read<K>(); }
    } else if constexpr (MODE == 2) {
        for (long long t = gw; t < ntiles; t += tw)
            if (lane == 0) bulk_store(out + tile_addr(t, tile_floats, pattern, plane_floats), s, tile_floats * 4);
    } else if constexpr (MODE == 3) {
        // tile t is patched once K newer tiles have been issued
        long long pend[K + 1];
        int head = 0, count = 0;
        for (long long t = gw; ; t += tw) {
            const bool live = t < ntiles;
            if (live) {
                if (lane == 0) bulk_store(out + tile_addr(t, tile_floats, pattern, plane_floats), s, tile_floats * 4);
                pend[(head + count) % (K + 1)] = t; ++count;
            }
            if (count > K || (!live && count > 0)) {
                if (lane == 0) { if (live) wait_full<K>(); else wait_full<0>(); }
                __syncwarp();
                const long long pt = pend[head]; head = (head + 1) % (K + 1); --count;
                float* dst = out + tile_addr(pt, tile_floats, pattern, plane_floats);
                for (int c = 0; c < cells; ++c) dst[(lane * 37 + c * 1031) % tile_floats] = (float)pt;
            }
            if (!live && count == 0) break;
        }
    }
    if (lane == 0) wait_full<0>();
}

int main()
{
    const long long bytes = 288LL * 1024 * 1024;
    float* buf[3];
    for (int i = 0; i < 3; i++) cudaMalloc(&buf[i], bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch(buf[i % 3]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; i++) launch(buf[i % 3]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-86s %6.1f us %5.0f GB/s (%s)\n", name, ms / 20 * 1e3, bytes / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    const long long plane = 32768;          // r = 32: 128 KB planes
#define RUN(MODE, K, label, smem_expr)                                                                                     \
    for (int pattern : {0, 1}) for (int tile_kb : {8, 32}) for (int warps : {1, 2, 4, 8, 16}) for (int ctas : {1, 2}) {      \
        if (pattern == 0 && tile_kb == 32 && MODE != 2) continue;                                                            \
        const int tf = tile_kb * 256; const long long nt = bytes / 4 / tf; const size_t sm = (smem_expr);                    \
        if (sm * ctas > 220 * 1024 || sm > 200 * 1024) continue;                                                             \
        if (MODE == 0 && (ctas == 2 || warps > 8)) continue;                                                                 \
        char nm[160]; snprintf(nm, 160, "%s pattern=%d tile=%dKB warps=%d ctas/SM=%d smem=%zuKB", label, pattern, tile_kb, warps, ctas, sm / 1024); \
        cudaFuncSetAttribute(zs_kernel<MODE, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                      \
        run(nm, [&](float* b) { zs_kernel<MODE, K><<<sms * ctas, warps * 32, sm>>>(b, nt, tf, 2, pattern, plane); });        \
    }
    RUN(0, 2, "ring 2 slots           ", (size_t)warps * 2 * tf * 4)
    RUN(2, 0, "const unpaced          ", (size_t)tf * 4)
    RUN(1, 0, "const paced read<=0    ", (size_t)tf * 4)
    RUN(1, 1, "const paced read<=1    ", (size_t)tf * 4)
    RUN(1, 3, "const paced read<=3    ", (size_t)tf * 4)
    RUN(3, 1, "const patch depth 1    ", (size_t)tf * 4)
    RUN(3, 2, "const patch depth 2    ", (size_t)tf * 4)
    RUN(3, 4, "const patch depth 4    ", (size_t)tf * 4)
    RUN(3, 8, "const patch depth 8    ", (size_t)tf * 4)
    RUN(3, 16, "const patch depth 16   ", (size_t)tf * 4)
    RUN(4, 0, "plain st.global.v4     ", (size_t)16)
    return 0;
}
