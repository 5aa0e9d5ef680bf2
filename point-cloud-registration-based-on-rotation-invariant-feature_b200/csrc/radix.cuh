// radix.cuh — stable LSD radix sort of (key, int payload) pairs, 8 bits per pass, sm_100a.  Used by the scan-sized voxelizer
// (32-bit (cloud, cell) keys, csrc/voxelize.cu) and by grid subsampling (64-bit cell keys, csrc/gridsub.cu) in the role the
// reference gives to float atomics / an unordered_map; it replaces the cub::DeviceRadixSort calls of round 1.
//
// Two launches per pass over tiles of 4096 pairs:
//   rs_hist     per tile: counts of the pass's 256 digit values                          -> hist[tile][256]
//   rs_scatter  per tile: (a) its global base per digit straight from the hist matrix (thread d sums column d over the
//               earlier tiles and over all tiles, then a block scan over the digits) — no separate scan launch;
//               (b) stable ranks: a warp owns 512 consecutive pairs and takes them 32 at a time, __match_any_sync groups the
//               lanes of equal digit, the group's lowest lane advances the warp's counter of that digit; an exclusive scan over
//               the 8 warps per digit completes the rank; (c) the pairs go to base + rank.
// Memory order (tile, warp, round, lane) is the order of the ranks, so equal keys keep their input order — the property the
// callers rely on (points of one cell summed in ascending point order).
//
// The number of passes is ceil(bits / 8).  `bits` is either known on the host (voxelizer) or only on the device (grid
// subsampling: the extent of the scan decides it); then every pass up to the host's bound is launched and the kernels of a pass
// that is not needed return at once, reading the largest key from device memory — no host round trip.  The sorted pairs end up
// in buffer (passes & 1): rs_result_in_b() tells host or device code which one.
#pragma once
#include "ri_common.cuh"

namespace ri_radix {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;       // 4096 pairs per CTA
constexpr int kWarps = kThreads / 32;

__host__ __device__ inline int passes_for_bits(int bits) { return (bits + 7) / 8; }
__device__ __forceinline__ int bits_of_max(unsigned long long maxkey) { return 64 - __clzll((long long)(maxkey | 1ull)); }
// device side: is the sorted result in the B buffers?  (maxkey_dev as given to the sort, host_bits its host-side bound)
__device__ __forceinline__ bool rs_result_in_b(const unsigned long long* maxkey_dev, int host_bits)
{
    int bits = host_bits;
    if (maxkey_dev != nullptr) { const int b = bits_of_max(*maxkey_dev); bits = b < bits ? b : bits; }
    return (passes_for_bits(bits) & 1) != 0;
}
__device__ __forceinline__ bool pass_needed(const unsigned long long* maxkey_dev, int shift)
{
    return maxkey_dev == nullptr || shift < bits_of_max(*maxkey_dev);
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
rs_hist_kernel(const KeyT* __restrict__ in, int n, int shift, const unsigned long long* __restrict__ maxkey_dev,
               int* __restrict__ hist)
{
    if (!pass_needed(maxkey_dev, shift)) return;
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kTile;
#pragma unroll 4
    for (int e = 0; e < kItems; ++e) {
        const int i = base + e * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&sh[(unsigned)(in[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    hist[(size_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
rs_scatter_kernel(const KeyT* __restrict__ in_k, const int* __restrict__ in_v, KeyT* __restrict__ out_k, int* __restrict__ out_v,
                  int n, int shift, int ntiles, const unsigned long long* __restrict__ maxkey_dev, const int* __restrict__ hist)
{
    if (!pass_needed(maxkey_dev, shift)) return;
    __shared__ int wcnt[kWarps][256];
    __shared__ int gbase[256];
    __shared__ int wtot[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) wcnt[q][tid] = 0;

    // (a) global base of digit d = tid for this tile
    int below = 0, total = 0;
    {
        const int* col = hist + tid;
        int t = 0;
        for (; t + 4 <= ntiles; t += 4) {
            const int v0 = col[(size_t)t * 256], v1 = col[(size_t)(t + 1) * 256], v2 = col[(size_t)(t + 2) * 256], v3 = col[(size_t)(t + 3) * 256];
            total += v0 + v1 + v2 + v3;
            below += (t < tile ? v0 : 0) + (t + 1 < tile ? v1 : 0) + (t + 2 < tile ? v2 : 0) + (t + 3 < tile ? v3 : 0);
        }
        for (; t < ntiles; ++t) {
            const int v = col[(size_t)t * 256];
            total += v;
            below += t < tile ? v : 0;
        }
    }
    int incl = total;                                        // inclusive scan over the 256 digits
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) woff += q < w ? wtot[q] : 0;
    gbase[tid] = woff + incl - total + below;
    __syncthreads();

    // (b) ranks inside the tile
    const int seg = tile * kTile + w * (32 * kItems);
    KeyT key[kItems];
    int val[kItems], rank[kItems];
    unsigned dig[kItems];
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const int i = seg + e * 32 + lane;
        const bool live = i < n;
        key[e] = live ? in_k[i] : (KeyT)0;
        val[e] = live ? in_v[i] : 0;
        dig[e] = live ? ((unsigned)(key[e] >> shift) & 255u) : (256u + (unsigned)lane);    // dead lanes match nobody
    }
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[e]);
        const int leader = __ffs(peers) - 1;
        const int before = __popc(peers & ((1u << lane) - 1u));
        int basec = 0;
        if (lane == leader && dig[e] < 256u) { basec = wcnt[w][dig[e]]; wcnt[w][dig[e]] = basec + __popc(peers); }
        basec = __shfl_sync(0xffffffffu, basec, leader);
        rank[e] = basec + before;
        __syncwarp();
    }
    __syncthreads();
    {                                                        // exclusive scan over the warps, per digit
        int run = 0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) { const int c = wcnt[q][tid]; wcnt[q][tid] = run; run += c; }
    }
    __syncthreads();

    // (c) scatter
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
        if (dig[e] < 256u) {
            const int pos = gbase[dig[e]] + wcnt[w][dig[e]] + rank[e];
            out_k[pos] = key[e];
            out_v[pos] = val[e];
        }
    }
}

inline size_t hist_bytes(long long n) { return (size_t)((n + kTile - 1) / kTile) * 256 * sizeof(int); }

// Sorts n pairs by the low `host_bits` bits of the key (device side: by the bits of *maxkey_dev if that is fewer).  The pairs
// start in (ka, va); (kb, vb) is the ping-pong partner; hist must hold hist_bytes(n).  Returns RI_OK or a CUDA error code.
template <typename KeyT>
inline int sort_pairs(KeyT* ka, KeyT* kb, int* va, int* vb, int n, int host_bits, const unsigned long long* maxkey_dev,
                      int* hist, cudaStream_t st)
{
    if (n <= 0) return RI_OK;
    const int ntiles = (n + kTile - 1) / kTile;
    const int passes = passes_for_bits(host_bits);
    for (int p = 0; p < passes; ++p) {
        KeyT* ik = (p & 1) ? kb : ka; KeyT* ok = (p & 1) ? ka : kb;      // executed passes are 0 .. passes-1 without gaps
        int* iv = (p & 1) ? vb : va; int* ov = (p & 1) ? va : vb;
        rs_hist_kernel<KeyT><<<ntiles, kThreads, 0, st>>>(ik, n, 8 * p, maxkey_dev, hist);
        rs_scatter_kernel<KeyT><<<ntiles, kThreads, 0, st>>>(ik, iv, ok, ov, n, 8 * p, ntiles, maxkey_dev, hist);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RI_OK : (int)e;
}

}  // namespace ri_radix
