// voxelize.cu — spherical and cube scatter-mean voxelization, sm_100a.
//
// Replaces  spherical_grid_stats_kernel + spherical_avg_voxelize_kernel
//           (/root/reference/PVCNN/modules/functional/src/spherical_voxelization/spherical_vox.cu:19-77, 91-125;
//            host: spherical_vox.cpp:17-46)
//      and  grid_stats_kernel + avg_voxelize_kernel
//           (.../voxelization/vox.cu:18-35, 49-73; host: vox.cpp:17-43).
// Outputs as the reference: out [B,C,s] fp32 (s = r^3), ind [B,N] i32 (-1 = undefined), cnt [B,s] i32.
//   ind, cnt : bit-exact (index arithmetic in ri_common.cuh::ri_sph_cell, counts are integers);
//   out      : sum_i feat[c,i] * (1.0f/cnt) exactly as the reference forms it, but summed in ascending point
//              order (deterministic) instead of float-atomic arrival order.
//
// The reference zero-fills out/ind/cnt (three memsets, 281 MB at B=32,C=67,r=32) and then issues C strided
// float atomics per point from one CTA per cloud.  ~98 % of the dense grid is zeros and the consumer is a dense
// Conv3d, so the compulsory traffic is "write the grid once".  Design here ("tile-composed writer"):
//
//   K1  vox_prepare   one CTA per cloud: bin every point (bit-exact), sort (cell,point) keys with a shared-
//                     memory bitonic network, emit the compact occupied-cell table of the cloud
//                     (cell id, first point, point count) and, per tile of the grid, where its cells start.
//   K2  vox_means     grid (8-channel groups, clouds): the mean of every occupied cell in every channel, one
//                     thread per cell, points summed in ascending order; written as a compact table
//                     means[b][c][cell slot] (coalesced).  The same CTA then emits the DGCNN edge features
//                     edge[b] = cat(feat - mean_of_own_cell, feat) point by point (fully coalesced) when asked to.
//   K3  vox_fill      persistent CTAs (2 per SM).  A work item is (cloud, grid tile, group of channel planes).
//                     Each CTA keeps a ring of zeroed 32 KB tiles in shared memory; per plane it patches the
//                     few occupied cells of the tile with their means (coalesced reads of the compact table,
//                     8 planes prefetched at a time) and ships the tile to HBM with ONE bulk async copy
//                     (cp.async.bulk shared->global, the TMA engine, SASS UBLKCP), overlapping the next plane's
//                     patching with the store.  The integer count grid goes out the same way as plane C.
//                     Nothing is memset, nothing is read back, `out` is written exactly once, no atomics at all.
//
//   Fallback (N > 4096 points per cloud, r^3 not a multiple of 4, or misaligned pointers): memset + integer
//   atomics for counts + float atomics for the means over a (point tile, channel group, cloud) grid.
#include "ri_common.cuh"
#include "radix.cuh"
#include "prologue_math.cuh"
#include <stdlib.h>

extern "C" int ri_voxel_edge_gather_f32(const float* avg, const float* feat, const int* inds, int B, int C, int N, int s,
                                        float* out, void* stream);

namespace {

constexpr int kPrepThreads = 512;
constexpr int kSmallCloudMax = 4096;     // K1 sorts a whole cloud inside one CTA
constexpr int kTileCells = 2048;         // 8 KB of fp32 per grid tile: one bulk store
constexpr int kFwWarps = 4;              // autonomous writer warps per CTA (grid writer), one CTA per SM (see the launch)
constexpr int kFwMaxWarps = 8;
constexpr int kFwSlots = 2;              // 8 KB tiles in a warp's ring
constexpr int kFillCtasPerSm = 1;
constexpr int kMaxFillCalls = 64;        // distinct cloud ranges one workspace can serve between two prepares
constexpr int kPlaneBatch = 4;            // channel planes whose table reads are issued together
constexpr int kPlaneGroup = 8;            // planes per work item (>= kFwSlots)
constexpr int kListCapDefault = 768;     // occupied cells of a tile whose ids / values a writer warp stages in shared memory
constexpr unsigned kNoCell = 0xffffffffu;

struct VoxWs {                            // per-cloud int32 workspace layout
    int stride, off_pid, off_cell, off_start, off_tile, off_meta, off_segof;
};
__host__ __device__ inline VoxWs vox_ws_layout(int N, int ntiles)
{
    VoxWs w;
    w.off_pid = 0;
    w.off_cell = N;
    w.off_start = 2 * N;
    w.off_tile = 3 * N + 1;
    w.off_meta = w.off_tile + ntiles + 1;
    w.off_segof = w.off_meta + 2;            // table slot of every point (-1: outside the grid)
    w.stride = (w.off_segof + N + 3) / 4 * 4;
    return w;
}

// ------------------------------------------------------------------------------------------------ K1
// coords: float [B,3,N] (SPH) or int [B,3,N] (cube).  P = power of two >= N.
template <bool SPH>
__global__ void __launch_bounds__(kPrepThreads)
vox_prepare_kernel(const void* __restrict__ coords_v, int N, int P, int r, int s, int tile_cells, int ntiles,
                   int* __restrict__ ind, int* __restrict__ ws, int* __restrict__ fill_counters)
{
    extern __shared__ unsigned long long skeys[];          // [P] keys, then [P] ints of segment cells
    if (blockIdx.x == 0 && threadIdx.x < 2 * kMaxFillCalls) fill_counters[threadIdx.x] = 0;   // the grid writer's work counters
    int* scell = reinterpret_cast<int*>(skeys + P);
    __shared__ int swarp_heads[kPrepThreads / 32];
    __shared__ int swarp_valid[kPrepThreads / 32];
    __shared__ int stotal[2];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const VoxWs L = vox_ws_layout(N, ntiles);
    int* W = ws + (size_t)b * L.stride;

    for (int i = tid; i < P; i += kPrepThreads) {
        unsigned long long key = ~0ull;
        if (i < N) {
            int cell;
            if (SPH) {
                const float* X = reinterpret_cast<const float*>(coords_v) + (size_t)b * 3 * N;
                cell = ri_sph_cell(X[i], X[i + N], X[i + 2 * (size_t)N], r);
            } else {
                const int* X = reinterpret_cast<const int*>(coords_v) + (size_t)b * 3 * N;
                cell = X[i] * r * r + X[i + N] * r + X[i + 2 * (size_t)N];          // vox.cu:31
            }
            ind[(size_t)b * N + i] = cell;
            // points outside the grid sort behind every occupied cell but keep their id (the fused edge
            // output still has to be written for them)
            const unsigned hi = (cell >= 0 && cell < s) ? (unsigned)cell : kNoCell;
            key = ((unsigned long long)hi << 32) | (unsigned)i;
        }
        skeys[i] = key;
    }
    __syncthreads();

    // bitonic sort, ascending: (cell, point) — points of one cell end up in ascending point order
    for (int size = 2; size <= P; size <<= 1) {
        for (int j = size >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += kPrepThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = skeys[i], c = skeys[l];
                const bool up = (i & size) == 0;
                if ((a > c) == up) { skeys[i] = c; skeys[l] = a; }
            }
            __syncthreads();
        }
    }

    // segment heads: each thread owns a run of E consecutive sorted slots
    const int E = (P + kPrepThreads - 1) / kPrepThreads;
    const int u0 = tid * E;
    int heads = 0, valid = 0;
    for (int e = 0; e < E; ++e) {
        const int u = u0 + e;
        if (u < P) {
            const unsigned long long key = skeys[u];
            if ((unsigned)(key >> 32) != kNoCell) {
                ++valid;
                if (u == 0 || (unsigned)(skeys[u - 1] >> 32) != (unsigned)(key >> 32)) ++heads;
            }
        }
    }
    // block-wide exclusive scan of `heads`, reduction of `valid`
    const int lane = tid & 31, wid = tid >> 5;
    int incl = heads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int vsum = valid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    if (lane == 31) swarp_heads[wid] = incl;
    if (lane == 0) swarp_valid[wid] = vsum;
    __syncthreads();
    if (wid == 0) {
        int h = lane < kPrepThreads / 32 ? swarp_heads[lane] : 0;
        int v = lane < kPrepThreads / 32 ? swarp_valid[lane] : 0;
        int hi = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, hi, o);
            if (lane >= o) hi += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < kPrepThreads / 32) swarp_heads[lane] = hi - h;    // exclusive warp offsets
        if (lane == kPrepThreads / 32 - 1) stotal[0] = hi;           // U = number of occupied cells
        if (lane == 0) stotal[1] = v;                                // points that landed in a cell
    }
    __syncthreads();
    int seg = swarp_heads[wid] + (incl - heads);
    const int U = stotal[0], nvalid = stotal[1];
    for (int e = 0; e < E; ++e) {
        const int u = u0 + e;
        if (u < P) {
            const unsigned long long key = skeys[u];
            if (u < N) W[L.off_pid + u] = (int)(unsigned)(key & 0xffffffffu);
            if ((unsigned)(key >> 32) != kNoCell) {
                const unsigned cell = (unsigned)(key >> 32);
                if (u == 0 || (unsigned)(skeys[u - 1] >> 32) != cell) {
                    W[L.off_cell + seg] = (int)cell;
                    W[L.off_start + seg] = u;
                    scell[seg] = (int)cell;
                    ++seg;
                }
                W[L.off_segof + (int)(unsigned)(key & 0xffffffffu)] = seg - 1;
            } else if (u < N) {
                W[L.off_segof + (int)(unsigned)(key & 0xffffffffu)] = -1;
            }
        }
    }
    if (tid == 0) { W[L.off_start + U] = nvalid; W[L.off_meta] = U; W[L.off_meta + 1] = nvalid; }
    __syncthreads();
    // first occupied-cell slot of every grid tile (lower bound on the sorted cell list)
    for (int t = tid; t <= ntiles; t += kPrepThreads) {
        const long long want = (long long)t * tile_cells;
        int lo = 0, hi = U;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((long long)scell[mid] < want) lo = mid + 1; else hi = mid;
        }
        W[L.off_tile + t] = lo;
    }
}

// ------------------------------------------------------------------------------------------------ K2
// means[b][c][slot] = sum over the cell's points (ascending point order) of feat[b][c][i] * (1.0f / count)
// (vox.cu:66-70 / spherical_vox.cu:112-116 form the same products; the reference adds them in atomic arrival order).
// With `edge` != null the CTA then writes the DGCNN edge features of its channels (pvconv.py:68-90):
//   edge[b, c, i] = feat - mean of i's cell (0 for a point outside the grid),  edge[b, C + c, i] = feat.
constexpr int kMeanChans = 8;
constexpr int kMeanThreads = 256;

__global__ void __launch_bounds__(kMeanThreads)
vox_means_kernel(const float* __restrict__ feat, const int* __restrict__ ws, int b0, int C, int N, int ntiles, int ucap,
                 int smem_means, float* __restrict__ means, float* __restrict__ edge)
{
    // The CTA's 8 feature rows are staged in shared memory (coalesced read, 32 KB at N = 1024): the per-cell gathers
    // then never touch L1, so the kernel is indifferent to the L1/shared split and can share SMs with the k-NN branch
    // under the grid writer's max-shared carveout.  Row stride N + 1 words spreads the 8 rows over the banks.
    extern __shared__ float sfeat[];                   // [kMeanChans][N + 1], then the cell means [kMeanChans][U]
    const int b = b0 + blockIdx.y, c0 = blockIdx.x * kMeanChans, tid = threadIdx.x;
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int* W = ws + (size_t)b * L.stride;
    const int* pid = W + L.off_pid;
    const int U = __ldg(W + L.off_meta);
    const int nch = min(kMeanChans, C - c0);
    const int ld = N + 1;
    const float* F = feat + ((size_t)b * C + c0) * N;
    float* M = means + ((size_t)b * C + c0) * ucap;
    float* smean = sfeat + kMeanChans * ld;

    for (int e = tid; e < nch * N; e += kMeanThreads) {
        const int j = e / N, i = e - j * N;
        sfeat[j * ld + i] = __ldg(F + (size_t)j * N + i);
    }
    __syncthreads();
    for (int sg = tid; sg < U; sg += kMeanThreads) {
        const int st = __ldg(W + L.off_start + sg);
        const int cnt = __ldg(W + L.off_start + sg + 1) - st;
        const float inv = __fdiv_rn(1.0f, (float)cnt);                                     // vox.cu:66
        float acc[kMeanChans];
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j) acc[j] = 0.f;
        for (int u = st; u < st + cnt; ++u) {                                               // ascending point order
            const int i = __ldg(pid + u);
#pragma unroll
            for (int j = 0; j < kMeanChans; ++j)
                if (j < nch) acc[j] = __fadd_rn(acc[j], __fmul_rn(sfeat[j * ld + i], inv));           // vox.cu:68-70
        }
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j)
            if (j < nch) {
                M[(size_t)j * ucap + sg] = acc[j];
                if (smem_means) smean[j * ucap + sg] = acc[j];
            }
    }
    if (edge == nullptr) return;
    __syncthreads();
    const int* segof = W + L.off_segof;
    float* E = edge + ((size_t)b * 2 * C + c0) * N;
    for (int i = tid; i < N; i += kMeanThreads) {
        const int sg = __ldg(segof + i);
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j)
            if (j < nch) {
                const float f = sfeat[j * ld + i];
                float mu = 0.f;
                if (sg >= 0) mu = smem_means ? smean[j * ucap + sg] : __ldcg(M + (size_t)j * ucap + sg);
                E[(size_t)j * N + i] = sg >= 0 ? __fsub_rn(f, mu) : 0.f;
                E[((size_t)C + j) * N + i] = f;
            }
    }
}

// ------------------------------------------------------------------------------------------ K0+K1+K2 fused
// The whole prefix of the voxel branch in ONE launch, for clouds that fit a CTA comfortably (N <= kFrontMaxN):
// coordinate prologue (prologue.cu) + prepare (K1) + means/edge (K2).  Grid = (8-channel groups, clouds); every CTA of a
// cloud repeats the cheap per-cloud part (normalise, bin, sort 1024 keys, build the cell table — in shared memory) and
// then does its own channel group; the CTA of group 0 also publishes norm_coords / vox_coords / ind and the tables the
// grid writer needs.  Three dependent, latency-bound launches (5 + 15 + 25 us, each with 32-288 CTAs) become one whose
// feature staging (cp.async) overlaps the sort.
constexpr int kFrontThreads = 512;
constexpr int kFrontMaxN = 1024;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(ri_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <bool SPH>
__global__ void __launch_bounds__(kFrontThreads)
vox_front_kernel(const float* __restrict__ points, int pstride, float* __restrict__ mean, int own_mean, float mean_factor,
                 const float* __restrict__ feat, int C, int N, int P, int r, int s, int tile_cells, int ntiles,
                 int shape, float eps, int norm_mode, int ucap,
                 float* __restrict__ norm_coords, int* __restrict__ vox_coords, int* __restrict__ ind,
                 int* __restrict__ ws, float* __restrict__ means, float* __restrict__ edge)
{
    extern __shared__ __align__(16) unsigned long long skeys[];   // [P] sort keys
    int* scell = reinterpret_cast<int*>(skeys + P);        // [P] cell of each table slot
    int* spid = scell + P;                                 // [N] cell-sorted point ids
    int* sstart = spid + N;                                // [N + 1] first sorted slot of each table slot
    int* ssegof = sstart + N + 1;                          // [N] table slot of each point (-1 outside the grid)
    float* sfeat = reinterpret_cast<float*>(ssegof + N + ((4 - ((3 * N + 1) & 3)) & 3));   // [kMeanChans][N + 4], 16-byte aligned
    const int ld = N + 4;
    float* smean = sfeat + kMeanChans * ld;                // [kMeanChans][ucap]
    __shared__ float sred[kFrontThreads / 32];
    __shared__ int swarp_heads[kFrontThreads / 32];
    __shared__ int swarp_valid[kFrontThreads / 32];
    __shared__ int stotal[2];

    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x;
    const bool publish = g == 0;
    const int c0 = g * kMeanChans;
    const int nch = min(kMeanChans, C - c0);
    const VoxWs L = vox_ws_layout(N, ntiles);
    int* W = ws + (size_t)b * L.stride;
    if (b == 0 && g == 0 && tid < 2 * kMaxFillCalls)            // the grid writer's work counters: behind the means table
        reinterpret_cast<int*>(means + (size_t)gridDim.y * C * ucap)[tid] = 0;

    // ---- this CTA's feature rows start flowing into shared memory now; they are first needed after the sort
    {
        // 16 bytes per cp.async where the rows allow it: as 4-byte copies (16 LDGSTS per thread) the issue loop itself held half
        // of the kernel's stall samples (ncu, round 2: the load/store unit's queue throttles the instructions behind it)
        const float* F = feat + ((size_t)b * C + c0) * N;
        if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(F) & 15) == 0) {
            const int n4 = N >> 2;
            for (int e = tid; e < nch * n4; e += kFrontThreads) {
                const int j = e / n4, i = (e - j * n4) << 2;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(ri_smem_u32(sfeat + j * ld + i)), "l"(F + (size_t)j * N + i) : "memory");
            }
        } else {
            for (int e = tid; e < nch * N; e += kFrontThreads) {
                const int j = e / N, i = e - j * N;
                cp_async4(sfeat + j * ld + i, F + (size_t)j * N + i);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- coordinate prologue (prologue.cu::prologue_kernel, same arithmetic)
    const float* Pt = points + (size_t)b * pstride * N;
    __shared__ float smean3[3];
    if (own_mean) {
        // The per-cloud mean in the order torch's reduce kernel uses for a contiguous row of 128 <= N < 8192 floats, N % 4 == 0
        // (ATen/native/cuda/Reduce.cuh: one warp per output, vectorize_input with four accumulators per lane over float4
        // loads strided by 32 lanes, ((v0 + v1) + v2) + v3, shuffle-down tree with offsets 16, 8, 4, 2, 1, times factor):
        // the same bits as coords.mean(2) — the engine verifies that once against torch before it relies on it.
        if (tid < 96) {
            const int a = tid >> 5, l = tid & 31;
            const float4* row = reinterpret_cast<const float4*>(Pt + (size_t)a * N);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            for (int idx = l; idx * 4 + 3 < N; idx += 32) {
                const float4 x = row[idx];
                v0 = __fadd_rn(v0, x.x); v1 = __fadd_rn(v1, x.y); v2 = __fadd_rn(v2, x.z); v3 = __fadd_rn(v3, x.w);
            }
            float v = __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), v3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
            if (l == 0) { smean3[a] = __fmul_rn(v, mean_factor); if (publish) mean[b * 3 + a] = smean3[a]; }
        }
        __syncthreads();
    } else if (tid < 3) {
        smean3[tid] = mean[b * 3 + tid];
    }
    if (!own_mean) __syncthreads();
    const float mx = smean3[0], my = smean3[1], mz = smean3[2];
    const size_t o3 = (size_t)b * 3 * N;
    float denom = 1.0f;
    if (shape != 0) {
        float m = 0.f;
        for (int i = tid; i < N; i += kFrontThreads)
            m = fmaxf(m, radius3(__fsub_rn(Pt[i], mx), __fsub_rn(Pt[i + N], my), __fsub_rn(Pt[i + 2 * (size_t)N], mz), norm_mode & 0xff));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((tid & 31) == 0) sred[tid >> 5] = m;
        __syncthreads();
        m = sred[0];
        for (int w = 1; w < kFrontThreads / 32; ++w) m = fmaxf(m, sred[w]);
        denom = (shape == 1) ? __fadd_rn(__fmul_rn(m, 2.0f), eps) : __fadd_rn(m, 1e-20f);
    }
    const float rf = (float)r, hi = (float)(r - 1);
    for (int i = tid; i < P; i += kFrontThreads) {
        unsigned long long key = ~0ull;
        if (i < N) {
            float v[3]; int vi[3] = {0, 0, 0};
            v[0] = ri_prologue_coord(__fsub_rn(Pt[i], mx), shape, denom, rf, hi, &vi[0]);
            v[1] = ri_prologue_coord(__fsub_rn(Pt[i + N], my), shape, denom, rf, hi, &vi[1]);
            v[2] = ri_prologue_coord(__fsub_rn(Pt[i + 2 * (size_t)N], mz), shape, denom, rf, hi, &vi[2]);
            int cell;
            if (SPH) cell = ri_sph_cell(v[0], v[1], v[2], r);
            else cell = vi[0] * r * r + vi[1] * r + vi[2];                             // vox.cu:31
            if (publish) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    norm_coords[o3 + (size_t)a * N + i] = v[a];
                    if (!SPH) vox_coords[o3 + (size_t)a * N + i] = vi[a];
                }
                ind[(size_t)b * N + i] = cell;
            }
            const unsigned hi32 = (cell >= 0 && cell < s) ? (unsigned)cell : kNoCell;
            key = ((unsigned long long)hi32 << 32) | (unsigned)i;
        }
        skeys[i] = key;
    }
    __syncthreads();

    // ---- K1: bitonic sort of (cell, point), cell table (vox_prepare_kernel, tables kept in shared memory)
    if (P == 2 * kFrontThreads && s < (1 << 21)) {
        // 1024 keys, two per thread (elements 2 tid, 2 tid + 1), packed into 32 bits (cell << 10 | point; 0x3FFFFE = outside the
        // grid, the same order as the 64-bit keys): the exchange distances 1 ... 32 stay inside a thread / a warp (min / max in
        // registers, one shuffle per key), only the 10 steps with distance >= 64 go through shared memory (two 4 KB buffers
        // alternate: one barrier per step).  The all-shared-memory network below took 55 barriers and 40 % of the kernel's
        // instructions (ncu, round 2).
        unsigned* const buf0 = reinterpret_cast<unsigned*>(skeys);
        unsigned* const buf1 = buf0 + 2 * kFrontThreads;
        auto pack = [&](unsigned long long k) -> unsigned {
            if (k == ~0ull) return 0xFFFFFFFFu;                      // padding slot (N < 1024): behind everything
            const unsigned c = (unsigned)(k >> 32);
            return ((c == kNoCell ? 0x3FFFFEu : c) << 10) | ((unsigned)k & 1023u);
        };
        const unsigned long long in0 = skeys[2 * tid], in1 = skeys[2 * tid + 1];
        unsigned k0 = pack(in0), k1 = pack(in1);
        __syncthreads();                                             // every key is in registers: the buffers may be overwritten
        const int lane_ = tid & 31;
        int flip = 0;
        for (int size = 2; size <= 2 * kFrontThreads; size <<= 1) {
            const bool up = ((2 * tid) & size) == 0;
            int j = size >> 1;
            for (; j >= 64; j >>= 1) {
                unsigned* const buf = flip ? buf1 : buf0;
                buf[2 * tid] = k0; buf[2 * tid + 1] = k1;
                __syncthreads();
                const int pt = tid ^ (j >> 1);
                const unsigned o0 = buf[2 * pt], o1 = buf[2 * pt + 1];
                const bool take_min = (((2 * tid) & j) == 0) == up;
                k0 = take_min ? min(k0, o0) : max(k0, o0);
                k1 = take_min ? min(k1, o1) : max(k1, o1);
                flip ^= 1;
            }
            for (; j >= 2; j >>= 1) {
                const unsigned o0 = __shfl_xor_sync(0xffffffffu, k0, j >> 1), o1 = __shfl_xor_sync(0xffffffffu, k1, j >> 1);
                const bool take_min = ((lane_ & (j >> 1)) == 0) == up;
                k0 = take_min ? min(k0, o0) : max(k0, o0);
                k1 = take_min ? min(k1, o1) : max(k1, o1);
            }
            const unsigned lo = min(k0, k1), hi2 = max(k0, k1);
            k0 = up ? lo : hi2;
            k1 = up ? hi2 : lo;
        }
        __syncthreads();                                             // the last readers of the exchange buffers are done
        auto unpack = [&](unsigned k) -> unsigned long long {
            if (k == 0xFFFFFFFFu) return ~0ull;
            const unsigned c = k >> 10;
            return ((unsigned long long)(c == 0x3FFFFEu ? kNoCell : c) << 32) | (k & 1023u);
        };
        skeys[2 * tid] = unpack(k0);
        skeys[2 * tid + 1] = unpack(k1);
        __syncthreads();
    } else {
        for (int size = 2; size <= P; size <<= 1) {
            for (int j = size >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (P >> 1); t += kFrontThreads) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const unsigned long long a = skeys[i], c = skeys[l];
                    const bool up = (i & size) == 0;
                    if ((a > c) == up) { skeys[i] = c; skeys[l] = a; }
                }
                __syncthreads();
            }
        }
    }
    const int E = (P + kFrontThreads - 1) / kFrontThreads;
    const int u0 = tid * E;
    int heads = 0, valid = 0;
    for (int e = 0; e < E; ++e) {
        const int u = u0 + e;
        if (u < P) {
            const unsigned long long key = skeys[u];
            if ((unsigned)(key >> 32) != kNoCell) {
                ++valid;
                if (u == 0 || (unsigned)(skeys[u - 1] >> 32) != (unsigned)(key >> 32)) ++heads;
            }
        }
    }
    const int lane = tid & 31, wid = tid >> 5;
    int incl = heads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int vsum = valid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    if (lane == 31) swarp_heads[wid] = incl;
    if (lane == 0) swarp_valid[wid] = vsum;
    __syncthreads();
    if (wid == 0) {
        int h = lane < kFrontThreads / 32 ? swarp_heads[lane] : 0;
        int v = lane < kFrontThreads / 32 ? swarp_valid[lane] : 0;
        int hsum = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, hsum, o);
            if (lane >= o) hsum += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < kFrontThreads / 32) swarp_heads[lane] = hsum - h;
        if (lane == kFrontThreads / 32 - 1) stotal[0] = hsum;
        if (lane == 0) stotal[1] = v;
    }
    __syncthreads();
    int seg = swarp_heads[wid] + (incl - heads);
    const int U = stotal[0], nvalid = stotal[1];
    for (int e = 0; e < E; ++e) {
        const int u = u0 + e;
        if (u < P) {
            const unsigned long long key = skeys[u];
            const int pt = (int)(unsigned)(key & 0xffffffffu);
            if (u < N) { spid[u] = pt; if (publish) W[L.off_pid + u] = pt; }
            if ((unsigned)(key >> 32) != kNoCell) {
                const unsigned cell = (unsigned)(key >> 32);
                if (u == 0 || (unsigned)(skeys[u - 1] >> 32) != cell) {
                    scell[seg] = (int)cell;
                    sstart[seg] = u;
                    if (publish) { W[L.off_cell + seg] = (int)cell; W[L.off_start + seg] = u; }
                    ++seg;
                }
                ssegof[pt] = seg - 1;
                if (publish) W[L.off_segof + pt] = seg - 1;
            } else if (u < N) {
                ssegof[pt] = -1;
                if (publish) W[L.off_segof + pt] = -1;
            }
        }
    }
    if (tid == 0) {
        sstart[U] = nvalid;
        if (publish) { W[L.off_start + U] = nvalid; W[L.off_meta] = U; W[L.off_meta + 1] = nvalid; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (publish) {
        for (int t = tid; t <= ntiles; t += kFrontThreads) {
            const long long want = (long long)t * tile_cells;
            int lo = 0, hi2 = U;
            while (lo < hi2) {
                const int mid = (lo + hi2) >> 1;
                if ((long long)scell[mid] < want) lo = mid + 1; else hi2 = mid;
            }
            W[L.off_tile + t] = lo;
        }
    }
    if (nch <= 0) return;

    // ---- K2: cell means of this CTA's channels (ascending point order), then the edge features
    float* M = means + ((size_t)b * C + c0) * ucap;
    for (int sg = tid; sg < U; sg += kFrontThreads) {
        const int st = sstart[sg];
        const int cnt = sstart[sg + 1] - st;
        const float inv = __fdiv_rn(1.0f, (float)cnt);                                     // vox.cu:66
        float acc[kMeanChans];
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j) acc[j] = 0.f;
        for (int u = st; u < st + cnt; ++u) {
            const int i = spid[u];
#pragma unroll
            for (int j = 0; j < kMeanChans; ++j)
                if (j < nch) acc[j] = __fadd_rn(acc[j], __fmul_rn(sfeat[j * ld + i], inv));           // vox.cu:68-70
        }
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j)
            if (j < nch) { M[(size_t)j * ucap + sg] = acc[j]; smean[j * ucap + sg] = acc[j]; }
    }
    if (edge == nullptr) return;
    __syncthreads();
    // rel_only (norm_mode bit 9): edge is [B,C,N] and receives the feat - mean half alone; the other half of the DGCNN
    // edge tensor is the input itself, which a caller on the far side of a PCIe link already holds
    const bool rel_only = (norm_mode & 0x200) != 0;
    float* Eo = edge + ((size_t)b * (rel_only ? 1 : 2) * C + c0) * N;
    for (int i = tid; i < N; i += kFrontThreads) {                 // point outer, channel inner: no division, one table-slot load
        const int sg = ssegof[i];
#pragma unroll
        for (int j = 0; j < kMeanChans; ++j) {
            if (j < nch) {
                const float f = sfeat[j * ld + i];
                Eo[(size_t)j * N + i] = sg >= 0 ? __fsub_rn(f, smean[j * ucap + sg]) : 0.f;
                if (!rel_only) Eo[((size_t)C + j) * N + i] = f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3
// The dense grid [C, r^3] (+ the count grid as plane C) of every cloud, written exactly once, by WARPS THAT EACH RUN THEIR OWN
// PIPELINE: a warp owns kFwSlots shared-memory tiles of kTileCells floats (8 KB), zeroed once; per (cloud, tile, plane)
// it writes the tile's few occupied cells into the next slot (values from the compact means table), and one lane ships
// the slot with ONE cp.async.bulk shared -> global (TMA engine, SASS UBLKCP).  No CTA barrier anywhere: the slot is
// guarded by the issuing lane's own bulk-group counter (wait_group.read) and two __syncwarp().  The CTA-synchronous
// form this replaces (256 threads around a 3 x 32 KB ring, two __syncthreads per tile) reached 5.1 TB/s alone and fell
// to 3 TB/s with k-NN warps competing for issue slots on the same SM — every barrier waited for the slowest of 8 warps;
// autonomous warps stall only themselves, and a micro-benchmark of the bare pattern (tools/exp_fillwarp.cu) streams at
// 6.0-6.1 TB/s with as little as 4 warps x 2 x 8 KB per SM.
// Work items = (cloud, tile, group of kPlaneGroup planes) from one global counter (a warp that becomes resident late
// simply takes fewer).  A warp keeps the tile's cell list (offset in the tile, table slot, point count) in registers and
// prefetches the NEXT item's list while it ships the current one.  Slots are never re-zeroed wholesale: the cells of the
// previous item's tile are cleared lazily, slot by slot, right before each slot's first reuse.
struct FillItem { int bt, b, t, p0, p1, sA, sB; };      // (cloud, tile) id, cloud, tile, plane range, the tile's table slots

__device__ __forceinline__ void fill_item_decode(FillItem& it, int item, int total_items, int groups, int pgroup, int planes,
                                                 int ntiles, int b0, const int* __restrict__ ws, const VoxWs& L)
{
    it.bt = -1; it.b = it.t = it.p0 = it.p1 = it.sA = it.sB = 0;
    if (item >= total_items) return;
    it.bt = item / groups;
    it.p0 = (item - it.bt * groups) * pgroup;
    it.p1 = min(planes, it.p0 + pgroup);
    it.t = it.bt % ntiles;
    it.b = b0 + it.bt / ntiles;
    const int* W = ws + (size_t)it.b * L.stride;
    it.sA = __ldg(W + L.off_tile + it.t);
    it.sB = __ldg(W + L.off_tile + it.t + 1);
}

template <int SLOTS>
__global__ void __launch_bounds__(32 * kFwMaxWarps)
vox_fill_kernel(const float* __restrict__ means, const int* __restrict__ ws, int b0, int B, int C, int N, int s,
                int tile_cells, int ntiles, int ucap, int pgroup, int kListCap, int* __restrict__ work_counter,
                float* __restrict__ out, int* __restrict__ cnt)
{
    const int kFwWarps = blockDim.x >> 5;
    // per warp: SLOTS tiles of tile_cells floats | 3 cell-id lists (previous / current / next item) | 2 value rows (this / next plane)
    extern __shared__ __align__(128) float sring[];
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int per_warp = SLOTS * tile_cells + 5 * kListCap;
    float* ring = sring + (size_t)wid * per_warp;
    int* lists = reinterpret_cast<int*>(ring + SLOTS * tile_cells);      // [3][kListCap]
    float* rows = ring + SLOTS * tile_cells + 3 * kListCap;               // [2][kListCap]
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int planes = C + 1;                              // plane C is the integer count grid
    const int groups = (planes + pgroup - 1) / pgroup;
    const int total_items = B * ntiles * groups;
    const int nwarps = gridDim.x * kFwWarps;

    for (int i = lane; i < SLOTS * tile_cells / 4; i += 32) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // Everything a plane needs arrives in shared memory by cp.async — no registers, any number of cells in flight at once:
    // the cell ids of an item's tile once per item, the tile's row of the compact means table once per plane.
    auto list_fetch = [&](const FillItem& it, int* dst) {
        if (it.bt >= 0) {
            const int* src = ws + (size_t)it.b * L.stride + L.off_cell + it.sA;
            const int n = min(it.sB - it.sA, kListCap);
            for (int i = lane; i < n; i += 32) cp_async4(dst + i, src + i);
        }
    };
    auto row_fetch = [&](const FillItem& it, int p, float* dst) {
        if (it.bt >= 0 && p < it.p1) {
            const int n = min(it.sB - it.sA, kListCap);
            if (p < C) {
                const float* src = means + ((size_t)it.b * C + p) * ucap + it.sA;
                for (int i = lane; i < n; i += 32) cp_async4(dst + i, src + i);
            } else {                                       // the count plane: point counts from the table's start offsets
                const int* S = ws + (size_t)it.b * L.stride + L.off_start + it.sA;
                for (int i = lane; i < n; i += 32) dst[i] = __int_as_float(__ldg(S + i + 1) - __ldg(S + i));
            }
        }
    };
    auto commit = [] { asm volatile("cp.async.commit_group;" ::: "memory"); };

    // Work items = (cloud, tile, group of `pgroup` planes), plane group fastest.  A plane-tile is NOT a constant amount of
    // work — spherical grids put most of a cloud's cells into one or two tiles — so the items are drawn from a global
    // counter (cutting the work into equal static runs per warp measured 87 us against 59 on the spherical grid, and 4.7
    // against 6.8 TB/s at r = 64).  The first item of every warp is its own number: no round trip to the counter before
    // the first store; the counter hands out the items from nwarps on.
    const int id0 = blockIdx.x * kFwWarps + wid;
    int id1 = 0;
    if (lane == 0) id1 = nwarps + atomicAdd(work_counter, 1);
    FillItem cur, nxt;
    fill_item_decode(cur, id0, total_items, groups, pgroup, planes, ntiles, b0, ws, L);
    int lc = 0;                                             // current item's list; (lc + 2) % 3 = previous, (lc + 1) % 3 = next
    int rc = 0;                                             // value row of the plane about to be shipped
    list_fetch(cur, lists + lc * kListCap);
    row_fetch(cur, cur.p0, rows + rc * kListCap);
    commit();
    int next_id = __shfl_sync(0xffffffffu, id1, 0);

    int slot = 0;
    int stale = 0;                                          // slots that still carry the previous item's cells
    int old_n = 0, old_sA = 0, old_cell_lo = 0;
    const int* oldW = ws;

    while (cur.bt >= 0) {
        // two items ahead: the atomic's reply is not needed before the end of this item; one item ahead: its tile header is
        // requested now and needed at this item's last plane
        int after = 0;
        if (lane == 0) after = nwarps + atomicAdd(work_counter, 1);
        fill_item_decode(nxt, next_id, total_items, groups, pgroup, planes, ntiles, b0, ws, L);

        const int* W = ws + (size_t)cur.b * L.stride;
        const int cell_lo = cur.t * tile_cells;
        const int ncell = min(tile_cells, s - cell_lo);
        const int n = cur.sB - cur.sA;                      // occupied cells of this tile
        const int nl = min(n, kListCap);
        const int* clist = lists + lc * kListCap;
        const int* olist = lists + ((lc + 2) % 3) * kListCap;

        for (int p = cur.p0; p < cur.p1; ++p) {
            // request the NEXT plane's values (or the next item's list and first row) before touching this plane
            if (p + 1 < cur.p1) row_fetch(cur, p + 1, rows + (rc ^ 1) * kListCap);
            else { list_fetch(nxt, lists + ((lc + 1) % 3) * kListCap); row_fetch(nxt, nxt.p0, rows + (rc ^ 1) * kListCap); }
            commit();
            float* tile = ring + slot * tile_cells;
            if (lane == 0) ri_bulk_wait_read<SLOTS - 1>();        // the copy that last used this slot has left shared memory
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // this plane's values (and this item's list) have landed
            __syncwarp();
            if (stale > 0) {                                       // first reuse of this slot since the item switch
                for (int i = lane; i < old_n; i += 32)
                    tile[(i < kListCap ? olist[i] : __ldg(oldW + L.off_cell + old_sA + i)) - old_cell_lo] = 0.f;
                --stale;
                __syncwarp();                                      // an old cell of one lane may be a new cell of another
            }
            const float* row = rows + rc * kListCap;
            for (int i = lane; i < nl; i += 32) tile[clist[i] - cell_lo] = row[i];
            for (int i = kListCap + lane; i < n; i += 32) {       // more cells in one tile than the lists hold
                const float x = p < C ? __ldg(means + ((size_t)cur.b * C + p) * ucap + cur.sA + i)
                                      : __int_as_float(__ldg(W + L.off_start + cur.sA + i + 1) - __ldg(W + L.off_start + cur.sA + i));
                tile[__ldg(W + L.off_cell + cur.sA + i) - cell_lo] = x;
            }
            ri_fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                void* dst = (p < C) ? (void*)(out + ((size_t)cur.b * C + p) * s + cell_lo)
                                    : (void*)(cnt + (size_t)cur.b * s + cell_lo);
                ri_bulk_store(dst, tile, (uint32_t)ncell * 4u);
                ri_bulk_commit();
            }
            slot = (slot + 1 == SLOTS) ? 0 : slot + 1;
            rc ^= 1;
        }
        // item switch.  Slots not reused since the LAST switch still hold the cells of the item before this one (only an item
        // of fewer planes than slots leaves such slots): two generations cannot be tracked lazily — clean everything now.
        if (stale > 0) {
            if (lane == 0) ri_bulk_wait_read<0>();
            __syncwarp();
            for (int k2 = 0; k2 < SLOTS; ++k2) {
                float* tile = ring + k2 * tile_cells;
                for (int i = lane; i < old_n; i += 32)
                    tile[(i < kListCap ? olist[i] : __ldg(oldW + L.off_cell + old_sA + i)) - old_cell_lo] = 0.f;
                for (int i = lane; i < n; i += 32)
                    tile[(i < kListCap ? clist[i] : __ldg(W + L.off_cell + cur.sA + i)) - cell_lo] = 0.f;
            }
            __syncwarp();
            stale = 0; old_n = 0;
        } else if (nxt.bt != cur.bt) {
            old_n = n; old_sA = cur.sA; old_cell_lo = cell_lo; oldW = W;
            stale = SLOTS;
        }                                                        // same (cloud, tile) next: its planes overwrite the same cells
        lc = (lc + 1) % 3;                                        // next -> current -> previous
        cur = nxt;
        next_id = __shfl_sync(0xffffffffu, after, 0);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (lane == 0) {
        ri_bulk_wait<0>();
        // the last warp of the launch to run out of work resets the counters: the next launch on this workspace finds zeros
        if (atomicAdd(work_counter + 1, 1) == nwarps - 1) { work_counter[1] = 0; __threadfence(); work_counter[0] = 0; }
    }
}

// Measured and rejected in round 2 (all bit-identical to this kernel; tools/exp_zerostream.cu, profiles/r2_fill_experiments.md):
//  * "zero-stream": every bulk store sourced from ONE constant zero tile, the occupied cells patched afterwards by ordinary
//    4-byte stores once the issuing lane had seen the bulk group complete — 82-92 us against 59.  Zeros alone stream at
//    6.0 TB/s from a constant tile (paced or not, like plain st.global.v4), but the patches need the FULL completion of the
//    bulk group, and with more than ~16 bulk stores outstanding per SM the lines have left L2 when the patch arrives
//    (5.25 TB/s at depth 1 x 8 warps, 2.5 TB/s at depth >= 4).  Patching in shared memory BEFORE the store keeps DRAM at one
//    write per line: 5.9 TB/s for the bare pattern.
//  * "register lists": the tile's cell offsets and the plane's values kept in registers (values loaded a plane ahead into
//    alternating register sets) instead of cp.async-staged shared-memory lists — a third of the instructions per tile, yet
//    60-61 us with 8 warps and 64-83 us with 4-6 (185 registers; crowded spherical tiles overflow 16 cells per lane).
//  * small work items (2 planes) for the last 10-40 % of the (cloud, tile) pairs against the tail of the dynamic hand-out:
//    58.1 us against 60.8 alone on the spherical grid, nothing on the cube grid, nothing with two batches in flight.

// --------------------------------------------------------------------------------------- fallback path
template <bool SPH>
__global__ void __launch_bounds__(256)
vox_index_atomic_kernel(const void* __restrict__ coords_v, int N, int r, int s, int* __restrict__ ind, int* __restrict__ cnt)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    int cell;
    if (SPH) {
        const float* X = reinterpret_cast<const float*>(coords_v) + (size_t)b * 3 * N;
        cell = ri_sph_cell(X[i], X[i + N], X[i + 2 * (size_t)N], r);
    } else {
        const int* X = reinterpret_cast<const int*>(coords_v) + (size_t)b * 3 * N;
        cell = X[i] * r * r + X[i + N] * r + X[i + 2 * (size_t)N];
    }
    ind[(size_t)b * N + i] = cell;
    if (cell >= 0 && cell < s) atomicAdd(cnt + (size_t)b * s + cell, 1);
}

constexpr int kScatterChans = 8;
__global__ void __launch_bounds__(256)
vox_scatter_atomic_kernel(const float* __restrict__ feat, const int* __restrict__ ind, const int* __restrict__ cnt,
                          int C, int N, int s, float* __restrict__ out)
{
    const int b = blockIdx.z;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int pos = ind[(size_t)b * N + i];
    if (pos < 0 || pos >= s) return;
    const int n = cnt[(size_t)b * s + pos];
    if (n <= 0) return;
    const float inv = __fdiv_rn(1.0f, (float)n);
    const int c0 = blockIdx.y * kScatterChans, c1 = min(C, c0 + kScatterChans);
    for (int c = c0; c < c1; ++c)
        atomicAdd(out + ((size_t)b * C + c) * s + pos, __fmul_rn(feat[((size_t)b * C + c) * N + i], inv));
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct VoxPlan { bool tiled; int tile_cells, ntiles, P; VoxWs L; };
inline int vox_tile_cells()
{
    const int v = ri_env().fill_tile_cells;
    return (v >= 256 && v <= 16384 && v % 4 == 0) ? v : kTileCells;
}

VoxPlan vox_plan(int N, int r, const void* out, const void* cnt)
{
    VoxPlan p;
    const long long s = (long long)r * r * r;
    p.tiled = N <= kSmallCloudMax && (s % 4 == 0) &&
              ((uintptr_t)out % 16 == 0) && ((uintptr_t)cnt % 16 == 0);
    const int tc = vox_tile_cells();
    p.tile_cells = (int)(s < tc ? s : tc);
    p.ntiles = (int)((s + p.tile_cells - 1) / p.tile_cells);
    p.P = next_pow2(N < 2 ? 2 : N);
    p.L = vox_ws_layout(N, p.ntiles);
    return p;
}

// work counters of the grid writer: behind the compact means table
inline int* vox_fill_counters(int* ws, const VoxPlan& plan, int B, int C, int N)
{
    float* means = reinterpret_cast<float*>(ws + (size_t)B * plan.L.stride);
    return reinterpret_cast<int*>(means + (size_t)B * C * ((N + 3) / 4 * 4));
}

size_t vox_ws_need(const VoxPlan& plan, int B, int C, int N)
{
    return (size_t)B * plan.L.stride * sizeof(int) + (size_t)B * C * ((N + 3) / 4 * 4) * sizeof(float) +
           2 * kMaxFillCalls * sizeof(int);                // work counters of the fill launches
}

// K1 for all B clouds: ind + the per-cloud cell tables in the workspace
template <bool SPH>
int vox_prepare_launch(const void* coords, int B, int C, int N, int r, int s, const VoxPlan& plan, int* ind, int* ws, cudaStream_t st)
{
    const size_t smem1 = (size_t)plan.P * (sizeof(unsigned long long) + sizeof(int));
    RI_KERNEL_SETUP(vox_prepare_kernel<SPH>, true, ri_step_carveout_percent());
    vox_prepare_kernel<SPH><<<B, kPrepThreads, smem1, st>>>(coords, N, plan.P, r, s, plan.tile_cells, plan.ntiles, ind, ws,
                                                            vox_fill_counters(ws, plan, B, C, N));
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// K2 for the clouds [b0, b1) of a batch of B whose tables are in the workspace: the compact table of cell means and,
// when `edge` is given, the DGCNN edge features (whole-batch array).
int vox_means_launch(const float* feat, int B, int C, int N, int b0, int b1, const VoxPlan& plan,
                     float* edge, int* ws, cudaStream_t st)
{
    float* means = reinterpret_cast<float*>(ws + (size_t)B * plan.L.stride);
    const int ucap = (N + 3) / 4 * 4;
    const int nb = b1 - b0;
    if (nb <= 0 || C <= 0) return RI_OK;
    dim3 gm((C + kMeanChans - 1) / kMeanChans, nb);
    // feature rows always staged (<= 128 KB at N = 4096); the means copy only while the CTA stays under ~100 KB
    const int smem_means = (edge != nullptr && (size_t)kMeanChans * (N + 1 + ucap) * sizeof(float) <= 100 * 1024) ? 1 : 0;
    const size_t smem_m = (size_t)kMeanChans * (N + 1 + (smem_means ? ucap : 0)) * sizeof(float);
    RI_KERNEL_SETUP(vox_means_kernel, true, ri_step_carveout_percent());
    vox_means_kernel<<<gm, kMeanThreads, smem_m, st>>>(feat, ws, b0, C, N, plan.ntiles, ucap, smem_means, means, edge);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// K3 for the clouds [b0, b1): the dense grid and the count grid from the tables and the means in the workspace.
int vox_fill_launch(int B, int C, int N, int s, int b0, int b1, const VoxPlan& plan,
                    float* out, int* cnt, int* ws, cudaStream_t st)
{
    float* means = reinterpret_cast<float*>(ws + (size_t)B * plan.L.stride);
    const int ucap = (N + 3) / 4 * 4;
    const int nb = b1 - b0;
    if (nb <= 0) return RI_OK;
    // A pair of work counters (items drawn, warps finished) per fill launch, slot chosen by the first cloud.  They are
    // zeroed by the prepare step of the workspace and again by every fill launch when its last warp retires, so no memset
    // is enqueued here.  Launches that may overlap in time must not share a slot: keep the fills of one workspace on one
    // stream (or their b0 distinct modulo kMaxFillCalls).
    // Footprint: 4 warps x (2 x 8 KB tiles + 5 x 768 x 4 B of lists and rows) = 124 KB, 128 threads x 96 registers.
    // Measured (tools/tune_lanes.py, 32 x 1024, r = 32, two batches in flight; grid writer alone / step one at a time / in flight):
    //   4 warps, 768-entry lists (124 KB)   cu_dg 60.6 / 166 / 144.9 us   sph_dg 59.3 / 137 / 120.8 us      <- default
    //   3 warps, 768-entry lists ( 93 KB)   cu_dg 68.3 / 175 / 140.7 us   sph_dg 65.4 / 142 / 121 us
    //   4 warps, 1024-entry lists (144 KB)  cu_dg 60.5 / 166 / 155.4 us   sph_dg 59.4 / 136 / 128.6 us
    // (the smaller the writer, the more often it shares an SM with the streaming devoxelizer or the prefix kernel of the
    // other batch in flight); 512-entry lists send the crowded tiles of the spherical grid down the slow path (78 us alone).
    int* counter = vox_fill_counters(ws, plan, B, C, N) + 2 * (b0 % kMaxFillCalls);
    const RiEnv& env = ri_env();
    const int sms = ri_num_sms();
    const int ctas_per_sm = (env.fill_ctas >= 1 && env.fill_ctas <= 4) ? env.fill_ctas : kFillCtasPerSm;
    const int slots = (env.fill_ring >= 2 && env.fill_ring <= 4) ? env.fill_ring : kFwSlots;
    const int warps = (env.fill_warps >= 1 && env.fill_warps <= kFwMaxWarps) ? env.fill_warps : kFwWarps;
    const int listcap = (env.fill_listcap >= 32 && env.fill_listcap <= 4096) ? (env.fill_listcap + 3) / 4 * 4 : kListCapDefault;
    size_t smem2 = (size_t)warps * ((size_t)slots * plan.tile_cells + 5 * listcap) * sizeof(float);
    if (env.fill_pad_kb >= 0 && env.fill_pad_kb <= 64) smem2 += (size_t)env.fill_pad_kb << 10;
    auto kern = slots == 2 ? vox_fill_kernel<2> : slots == 3 ? vox_fill_kernel<3> : vox_fill_kernel<4>;
    RI_KERNEL_SETUP(kern, true, ri_step_carveout_percent());
    const int pgroup = (env.fill_group >= slots && env.fill_group <= 64) ? env.fill_group : kPlaneGroup;
    const long long items = (long long)nb * plan.ntiles * ((C + 1 + pgroup - 1) / pgroup);
    if (items > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    long long grid = (long long)ctas_per_sm * sms;
    const long long want = (items + warps - 1) / warps;
    if (grid > want) grid = want;
    kern<<<(unsigned)grid, 32 * warps, smem2, st>>>(means, ws, b0, nb, C, N, s, plan.tile_cells, plan.ntiles,
                                                    ucap, pgroup, listcap, counter, out, cnt);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// --------------------------------------------------------------------------------------- scan-sized clouds (N > 4096)
// The same three phases — cell-sorted point list + occupied-cell table, compact cell means in ascending point order, dense
// grid written once by vox_fill_kernel — for clouds that do not fit one CTA's shared-memory sort (ICL-NUIM-sized scans,
// BASELINE configs[3]).  The (cloud, cell) keys of the whole batch go through ONE stable radix sort (csrc/radix.cuh:
// ceil(bits / 8) passes) of 32-bit keys  b * (s + 1) + cell  (cell = s for points outside the grid), the tables are rebuilt per
// cloud by a block-wide head-flag scan, and the means are summed per (cell, channel) thread in sorted = ascending point
// order: deterministic and equal to the oracle bit for bit, where the atomic path's float atomicAdd order changes from
// run to run (the reference has the same property).

template <bool SPH>
__global__ void __launch_bounds__(256)
vox_keys_large_kernel(const void* __restrict__ coords_v, int N, int r, int s, int* __restrict__ ind,
                      unsigned* __restrict__ keys, int* __restrict__ vals)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    int cell;
    if (SPH) {
        const float* X = reinterpret_cast<const float*>(coords_v) + (size_t)b * 3 * N;
        cell = ri_sph_cell(X[i], X[i + N], X[i + 2 * (size_t)N], r);
    } else {
        const int* X = reinterpret_cast<const int*>(coords_v) + (size_t)b * 3 * N;
        cell = X[i] * r * r + X[i + N] * r + X[i + 2 * (size_t)N];                  // vox.cu:31
    }
    ind[(size_t)b * N + i] = cell;
    const unsigned c = (cell >= 0 && cell < s) ? (unsigned)cell : (unsigned)s;
    keys[(size_t)b * N + i] = (unsigned)b * (unsigned)(s + 1) + c;
    vals[(size_t)b * N + i] = i;
}

// sorted keys / point ids of a cloud -> the workspace tables of vox_ws_layout, in four small grid-wide launches (a single
// CTA per cloud took 131 us for 8 x 47k points): A per-chunk head / valid counts, B per-cloud scan of the chunk counts,
// C per-chunk table writes, D per-tile offsets.
constexpr int kLgChunk = 1024;                                   // sorted slots per CTA in passes A and C (one per thread)

__device__ __forceinline__ void lg_flags(const unsigned* K, unsigned kbase, int s, int N, int u, int& head, int& valid)
{
    head = 0; valid = 0;
    if (u < N) {
        const unsigned c = K[u] - kbase;
        if (c < (unsigned)s) { valid = 1; head = (u == 0 || K[u - 1] != K[u]) ? 1 : 0; }
    }
}

__global__ void __launch_bounds__(kLgChunk)
vox_table_count_kernel(const unsigned* __restrict__ keys, int N, int s, int nchunks, int* __restrict__ chunk_counts)
{
    __shared__ int sh[kLgChunk / 32], sv[kLgChunk / 32];
    const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int head, valid;
    lg_flags(keys + (size_t)b * N, (unsigned)b * (unsigned)(s + 1), s, N, ch * kLgChunk + tid, head, valid);
    const unsigned hb = __ballot_sync(0xffffffffu, head), vb = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) { sh[wid] = __popc(hb); sv[wid] = __popc(vb); }
    __syncthreads();
    if (tid == 0) {
        int h = 0, v = 0;
        for (int w = 0; w < kLgChunk / 32; ++w) { h += sh[w]; v += sv[w]; }
        chunk_counts[((size_t)b * nchunks + ch) * 2] = h;
        chunk_counts[((size_t)b * nchunks + ch) * 2 + 1] = v;
    }
}

__global__ void vox_table_scan_kernel(int N, int ntiles, int nchunks, int* __restrict__ chunk_counts, int* __restrict__ ws,
                                      int* __restrict__ fill_counters)
{
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < 2 * kMaxFillCalls; i += blockDim.x) fill_counters[i] = 0;
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    const VoxWs L = vox_ws_layout(N, ntiles);
    int* W = ws + (size_t)b * L.stride;
    int h = 0, v = 0;
    for (int ch = 0; ch < nchunks; ++ch) {                        // a few dozen chunks per cloud
        int* c = chunk_counts + ((size_t)b * nchunks + ch) * 2;
        const int hh = c[0];
        c[0] = h;                                                 // exclusive prefix of the heads
        h += hh; v += c[1];
    }
    W[L.off_meta] = h; W[L.off_meta + 1] = v; W[L.off_start + h] = v;
}

__global__ void __launch_bounds__(kLgChunk)
vox_table_write_kernel(const unsigned* __restrict__ keys, const int* __restrict__ vals, int N, int s, int ntiles, int nchunks,
                       const int* __restrict__ chunk_counts, int* __restrict__ ws)
{
    __shared__ int sh[kLgChunk / 32];
    const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const VoxWs L = vox_ws_layout(N, ntiles);
    int* W = ws + (size_t)b * L.stride;
    const unsigned* K = keys + (size_t)b * N;
    const int u = ch * kLgChunk + tid;
    int head, valid;
    lg_flags(K, (unsigned)b * (unsigned)(s + 1), s, N, u, head, valid);
    const unsigned hb = __ballot_sync(0xffffffffu, head);
    if (lane == 0) sh[wid] = __popc(hb);
    __syncthreads();
    int base = chunk_counts[((size_t)b * nchunks + ch) * 2];
    for (int w = 0; w < wid; ++w) base += sh[w];
    const int seg = base + __popc(hb & ((1u << lane) - 1)) + head - 1;    // table slot of this sorted position's cell
    if (u < N) {
        const int pt = vals[(size_t)b * N + u];
        W[L.off_pid + u] = pt;
        if (valid) {
            if (head) { W[L.off_cell + seg] = (int)(K[u] - (unsigned)b * (unsigned)(s + 1)); W[L.off_start + seg] = u; }
            W[L.off_segof + pt] = seg;
        } else {
            W[L.off_segof + pt] = -1;
        }
    }
}

__global__ void vox_table_tiles_kernel(int N, int tile_cells, int ntiles, int* __restrict__ ws)
{
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const VoxWs L = vox_ws_layout(N, ntiles);
    int* W = ws + (size_t)b * L.stride;
    const int U = W[L.off_meta];
    const long long want = (long long)t * tile_cells;
    int lo = 0, hi2 = U;
    while (lo < hi2) {
        const int mid = (lo + hi2) >> 1;
        if ((long long)W[L.off_cell + mid] < want) lo = mid + 1; else hi2 = mid;
    }
    W[L.off_tile + t] = lo;
}

// one thread per occupied cell and channel group: sum f * (1/n) over the cell's points in ascending point order (vox.cu:61-72)
constexpr int kLgChans = 4;
__global__ void __launch_bounds__(256)
vox_means_large_kernel(const float* __restrict__ feat, const int* __restrict__ ws, int C, int N, int ntiles, int ucap,
                       float* __restrict__ means)
{
    const int b = blockIdx.z;
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int* W = ws + (size_t)b * L.stride;
    const int U = W[L.off_meta];
    const int sg = blockIdx.x * 256 + threadIdx.x;
    if (sg >= U) return;
    const int c0 = blockIdx.y * kLgChans;
    const int st = W[L.off_start + sg], cnt = W[L.off_start + sg + 1] - st;
    const float inv = __fdiv_rn(1.0f, (float)cnt);                                   // vox.cu:66
    const float* F = feat + (size_t)b * C * N;
    float acc[kLgChans];
#pragma unroll
    for (int j = 0; j < kLgChans; ++j) acc[j] = 0.f;
    for (int u = st; u < st + cnt; ++u) {
        const int i = W[L.off_pid + u];
#pragma unroll
        for (int j = 0; j < kLgChans; ++j)
            if (c0 + j < C) acc[j] = __fadd_rn(acc[j], __fmul_rn(__ldg(F + (size_t)(c0 + j) * N + i), inv));
    }
    float* M = means + (size_t)b * C * ucap;
#pragma unroll
    for (int j = 0; j < kLgChans; ++j)
        if (c0 + j < C) M[(size_t)(c0 + j) * ucap + sg] = acc[j];
}

// edge [B,2C,N] from the compact means table (pvconv.py:68-90)
__global__ void __launch_bounds__(256)
vox_edge_large_kernel(const float* __restrict__ feat, const int* __restrict__ ws, const float* __restrict__ means,
                      int C, int N, int ntiles, int ucap, float* __restrict__ edge)
{
    const int b = blockIdx.z, c = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int sg = ws[(size_t)b * L.stride + L.off_segof + i];
    const float f = feat[((size_t)b * C + c) * N + i];
    float* E = edge + (size_t)b * 2 * C * N;
    E[(size_t)c * N + i] = sg >= 0 ? __fsub_rn(f, means[((size_t)b * C + c) * ucap + sg]) : 0.f;
    E[((size_t)C + c) * N + i] = f;
}

struct VoxLarge { size_t keys_in, keys_out, vals_in, vals_out, chunks, hist, total; int bits, nchunks; };

// layout of the extra workspace of the scan-sized path, placed after the tables / means / counters of vox_ws_need
static bool vox_large_layout(int B, int N, long long s, size_t base, VoxLarge& V)
{
    const long long nkeys = (long long)B * (s + 1);
    if (nkeys > 0xffffffffLL || (long long)B * N > 0x7fffffffLL) return false;
    int bits = 1;
    while (bits < 32 && (1ll << bits) < nkeys) ++bits;
    V.bits = bits;
    const size_t n = (size_t)B * N;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = al(base);
    V.keys_in = o; o += al(n * 4);
    V.keys_out = o; o += al(n * 4);
    V.vals_in = o; o += al(n * 4);
    V.vals_out = o; o += al(n * 4);
    V.nchunks = (N + kLgChunk - 1) / kLgChunk;
    V.chunks = o; o += al((size_t)B * V.nchunks * 2 * sizeof(int));
    V.hist = o; o += al(ri_radix::hist_bytes((long long)n));
    V.total = o;
    return true;
}

template <bool SPH>
int voxelize_impl(const float* feat, const void* coords, int B, int C, int N, int r,
                  float* out, int* ind, int* cnt, float* edge, void* workspace, size_t ws_bytes, cudaStream_t st)
{
    if (B < 0 || C < 0 || N < 0 || r <= 0 || r > 1024) return RI_ERR_BAD_ARG;
    const long long s_ll = (long long)r * r * r;
    if (s_ll > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    const int s = (int)s_ll;
    if (B == 0) return RI_OK;
    const VoxPlan plan = vox_plan(N, r, out, cnt);
    if (plan.tiled && N > 0) {
        if (workspace == nullptr || ws_bytes < vox_ws_need(plan, B, C, N)) return RI_ERR_WORKSPACE;
        int* ws = reinterpret_cast<int*>(workspace);
        const int rc = vox_prepare_launch<SPH>(coords, B, C, N, r, s, plan, ind, ws, st);
        if (rc != RI_OK) return rc;
        const int rc2 = vox_means_launch(feat, B, C, N, 0, B, plan, edge, ws, st);
        if (rc2 != RI_OK) return rc2;
        return vox_fill_launch(B, C, N, s, 0, B, plan, out, cnt, ws, st);
    }
    // scan-sized clouds: global sort + the same means / grid-writer phases (deterministic); needs the larger workspace
    if (N > kSmallCloudMax && (s_ll % 4 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)cnt % 16 == 0) && B <= 65535 &&
        C <= 65535 && workspace != nullptr && !ri_env().vox_atomic) {
        VoxLarge V;
        const size_t base = vox_ws_need(plan, B, C, N);
        if (vox_large_layout(B, N, s_ll, base, V) && ws_bytes >= V.total) {
            unsigned char* wb = reinterpret_cast<unsigned char*>(workspace);
            int* ws = reinterpret_cast<int*>(workspace);
            unsigned* keys_in = reinterpret_cast<unsigned*>(wb + V.keys_in);
            unsigned* keys_out = reinterpret_cast<unsigned*>(wb + V.keys_out);
            int* vals_in = reinterpret_cast<int*>(wb + V.vals_in);
            int* vals_out = reinterpret_cast<int*>(wb + V.vals_out);
            vox_keys_large_kernel<SPH><<<dim3((N + 255) / 256, B), 256, 0, st>>>(coords, N, r, s, ind, keys_in, vals_in);
            RI_LAUNCH_CHECK();
            // stable sort of the batch's (cloud, cell) keys: ceil(bits / 8) passes of this repo's radix sort (radix.cuh)
            const int rcs = ri_radix::sort_pairs<unsigned>(keys_in, keys_out, vals_in, vals_out, (int)((size_t)B * N), V.bits,
                                                           nullptr, reinterpret_cast<int*>(wb + V.hist), st);
            if (rcs != RI_OK) return rcs;
            if ((ri_radix::passes_for_bits(V.bits) & 1) == 0) {      // an even number of passes leaves the result in the first pair
                unsigned* tk = keys_in; keys_in = keys_out; keys_out = tk;
                int* tv = vals_in; vals_in = vals_out; vals_out = tv;
            }
            int* chunks = reinterpret_cast<int*>(wb + V.chunks);
            vox_table_count_kernel<<<dim3(V.nchunks, B), kLgChunk, 0, st>>>(keys_out, N, s, V.nchunks, chunks);
            vox_table_scan_kernel<<<B, 32, 0, st>>>(N, plan.ntiles, V.nchunks, chunks, ws, vox_fill_counters(ws, plan, B, C, N));
            vox_table_write_kernel<<<dim3(V.nchunks, B), kLgChunk, 0, st>>>(keys_out, vals_out, N, s, plan.ntiles, V.nchunks, chunks, ws);
            vox_table_tiles_kernel<<<dim3((plan.ntiles + 256) / 256, B), 256, 0, st>>>(N, plan.tile_cells, plan.ntiles, ws);
            RI_LAUNCH_CHECK();
            float* means = reinterpret_cast<float*>(ws + (size_t)B * plan.L.stride);
            const int ucap = (N + 3) / 4 * 4;
            if (C > 0) {
                vox_means_large_kernel<<<dim3((N + 255) / 256, (C + kLgChans - 1) / kLgChans, B), 256, 0, st>>>(
                    feat, ws, C, N, plan.ntiles, ucap, means);
                RI_LAUNCH_CHECK();
                if (edge != nullptr) {
                    vox_edge_large_kernel<<<dim3((N + 255) / 256, C, B), 256, 0, st>>>(feat, ws, means, C, N, plan.ntiles, ucap, edge);
                    RI_LAUNCH_CHECK();
                }
            }
            return vox_fill_launch(B, C, N, s, 0, B, plan, out, cnt, ws, st);
        }
    }
    // fallback: memset + atomics (also covers N == 0)
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)B * C * s * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(cnt, 0, (size_t)B * s * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (N == 0) return RI_OK;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    dim3 g1((N + 255) / 256, B);
    vox_index_atomic_kernel<SPH><<<g1, 256, 0, st>>>(coords, N, r, s, ind, cnt);
    RI_LAUNCH_CHECK();
    if (C > 0) {
        dim3 g2((N + 255) / 256, (C + kScatterChans - 1) / kScatterChans, B);
        vox_scatter_atomic_kernel<<<g2, 256, 0, st>>>(feat, ind, cnt, C, N, s, out);
        RI_LAUNCH_CHECK();
        if (edge != nullptr) return ri_voxel_edge_gather_f32(out, feat, ind, B, C, N, s, edge, (void*)st);
    }
    return RI_OK;
}

}  // namespace

extern "C" size_t ri_voxelize_workspace_bytes(int B, int C, int N, int r)
{
    if (B <= 0 || N <= 0 || r <= 0) return 16;
    if (C < 0) C = 0;
    const long long s = (long long)r * r * r;
    const int tc = vox_tile_cells();
    const int tile_cells = (int)(s < tc ? s : tc);
    const int ntiles = (int)((s + tile_cells - 1) / tile_cells);
    size_t need = (size_t)B * vox_ws_layout(N, ntiles).stride * sizeof(int) +     // per-cloud tables
                  (size_t)B * C * ((N + 3) / 4 * 4) * sizeof(float) +            // compact cell means [B][C][<=N]
                  2 * kMaxFillCalls * sizeof(int) + 16;                           // work counters
    if (N > kSmallCloudMax) {                                                     // scan-sized clouds: sort buffers
        VoxLarge V;
        if (vox_large_layout(B, N, s, need, V)) need = V.total + 16;
    }
    return need;
}

extern "C" int ri_sph_voxelize_f32(const float* feat, const float* coords, int B, int C, int N, int r,
                                   float* out, int* ind, int* cnt, void* workspace, size_t ws_bytes, void* stream)
{
    return voxelize_impl<true>(feat, coords, B, C, N, r, out, ind, cnt, nullptr, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ri_cube_voxelize_f32(const float* feat, const int* coords, int B, int C, int N, int r,
                                    float* out, int* ind, int* cnt, void* workspace, size_t ws_bytes, void* stream)
{
    return voxelize_impl<false>(feat, coords, B, C, N, r, out, ind, cnt, nullptr, workspace, ws_bytes, (cudaStream_t)stream);
}

// Fused forms: voxelize AND emit the DGCNN edge features edge [B,2C,N] of the same points in the same pass
// (the voxelizer already holds every point's cell mean), so PVConv's gather never re-reads the dense grid.
extern "C" int ri_sph_voxelize_edge_f32(const float* feat, const float* coords, int B, int C, int N, int r,
                                        float* out, int* ind, int* cnt, float* edge,
                                        void* workspace, size_t ws_bytes, void* stream)
{
    if (edge == nullptr) return RI_ERR_BAD_ARG;
    return voxelize_impl<true>(feat, coords, B, C, N, r, out, ind, cnt, edge, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ri_cube_voxelize_edge_f32(const float* feat, const int* coords, int B, int C, int N, int r,
                                         float* out, int* ind, int* cnt, float* edge,
                                         void* workspace, size_t ws_bytes, void* stream)
{
    if (edge == nullptr) return RI_ERR_BAD_ARG;
    return voxelize_impl<false>(feat, coords, B, C, N, r, out, ind, cnt, edge, workspace, ws_bytes, (cudaStream_t)stream);
}

// ---- two-phase form, for schedules that pipeline the grid through L2 ------------------------------------
// prepare: ind [B,N] + the cell tables of ALL clouds (one launch); means: the compact cell-mean table (+ the DGCNN edge
// features) and fill: the dense grid + count grid, both for the clouds [b0, b1) only.  A caller can then run  fill(chunk i) -> consumer(chunk i)  with chunks small enough that the
// consumer (a devoxelize, a Conv3d) still finds the chunk's grid in the 126 MB L2.  Same workspace as the one-shot
// calls; RI_ERR_UNSUPPORTED when the shape falls outside the tiled path (N > 4096, r^3 % 4 != 0, misaligned outputs):
// use the one-shot entry points then.
template <bool SPH>
static int prepare_entry(const void* coords, int B, int C, int N, int r, int* ind, void* workspace, size_t ws_bytes, void* stream)
{
    if (B < 0 || C < 0 || N <= 0 || r <= 0 || r > 1024) return RI_ERR_BAD_ARG;
    const long long s_ll = (long long)r * r * r;
    if (s_ll > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    if (B == 0) return RI_OK;
    const VoxPlan plan = vox_plan(N, r, nullptr, nullptr);
    if (!plan.tiled) return RI_ERR_UNSUPPORTED;
    if (workspace == nullptr || ws_bytes < vox_ws_need(plan, B, C, N)) return RI_ERR_WORKSPACE;
    return vox_prepare_launch<SPH>(coords, B, C, N, r, (int)s_ll, plan, ind, reinterpret_cast<int*>(workspace), (cudaStream_t)stream);
}

extern "C" int ri_sph_voxelize_prepare_f32(const float* coords, int B, int C, int N, int r, int* ind,
                                           void* workspace, size_t ws_bytes, void* stream)
{
    return prepare_entry<true>(coords, B, C, N, r, ind, workspace, ws_bytes, stream);
}

extern "C" int ri_cube_voxelize_prepare_f32(const int* coords, int B, int C, int N, int r, int* ind,
                                            void* workspace, size_t ws_bytes, void* stream)
{
    return prepare_entry<false>(coords, B, C, N, r, ind, workspace, ws_bytes, stream);
}

extern "C" int ri_voxelize_means_f32(const float* feat, int B, int C, int N, int r, int b0, int b1, float* edge,
                                     void* workspace, size_t ws_bytes, void* stream)
{
    if (B < 0 || C < 0 || N <= 0 || r <= 0 || r > 1024 || b0 < 0 || b1 > B || b0 > b1) return RI_ERR_BAD_ARG;
    const long long s_ll = (long long)r * r * r;
    if (s_ll > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    const VoxPlan plan = vox_plan(N, r, nullptr, nullptr);
    if (!plan.tiled) return RI_ERR_UNSUPPORTED;
    if (workspace == nullptr || ws_bytes < vox_ws_need(plan, B, C, N)) return RI_ERR_WORKSPACE;
    return vox_means_launch(feat, B, C, N, b0, b1, plan, edge, reinterpret_cast<int*>(workspace), (cudaStream_t)stream);
}

extern "C" int ri_voxelize_fill_f32(int B, int C, int N, int r, int b0, int b1, float* out, int* cnt,
                                    void* workspace, size_t ws_bytes, void* stream)
{
    if (B < 0 || C < 0 || N <= 0 || r <= 0 || r > 1024 || b0 < 0 || b1 > B || b0 > b1) return RI_ERR_BAD_ARG;
    const long long s_ll = (long long)r * r * r;
    if (s_ll > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    const VoxPlan plan = vox_plan(N, r, out, cnt);
    if (!plan.tiled) return RI_ERR_UNSUPPORTED;
    if (workspace == nullptr || ws_bytes < vox_ws_need(plan, B, C, N)) return RI_ERR_WORKSPACE;
    return vox_fill_launch(B, C, N, (int)s_ll, b0, b1, plan, out, cnt, reinterpret_cast<int*>(workspace),
                           (cudaStream_t)stream);
}

// ---- fused prefix of the voxel branch ----------------------------------------------------------------------------
// ri_vox_prologue_f32 + ri_{sph,cube}_voxelize_prepare_f32 + ri_voxelize_means_f32 for all B clouds in ONE launch
// (vox_front_kernel); ri_voxelize_fill_f32 completes the voxelization.  points [B,pstride,N] (pstride 3 or 6), mean [B,3]
// (the caller's coords.mean(2)), shape 0 cube normalize=False / 1 cube normalize=True / 2 spherical, as ri_vox_prologue_f32.
// Outputs norm_coords [B,3,N], vox_coords [B,3,N] (cube shapes), ind [B,N], edge [B,2C,N] (nullable) and the workspace
// tables.  N <= 1024 and the tiled-path conditions, else RI_ERR_UNSUPPORTED (call the three entry points instead).
extern "C" int ri_vox_front_f32(const float* points, int pstride, const float* mean, const float* feat,
                                int B, int C, int N, int r, int shape, float eps, int norm_mode,
                                float* norm_coords, int* vox_coords, int* ind, float* edge,
                                void* workspace, size_t ws_bytes, void* stream)
{
    if (B < 0 || C < 0 || N <= 0 || r <= 0 || r > 1024 || (pstride != 3 && pstride != 6) || shape < 0 || shape > 2)
        return RI_ERR_BAD_ARG;
    if (norm_coords == nullptr || ind == nullptr || (shape != 2 && vox_coords == nullptr)) return RI_ERR_BAD_ARG;
    const long long s_ll = (long long)r * r * r;
    if (s_ll > 0x7fffffffLL || B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0) return RI_OK;
    const VoxPlan plan = vox_plan(N, r, nullptr, nullptr);
    if (!plan.tiled || N > kFrontMaxN) return RI_ERR_UNSUPPORTED;
    if (workspace == nullptr || ws_bytes < vox_ws_need(plan, B, C, N)) return RI_ERR_WORKSPACE;
    int* ws = reinterpret_cast<int*>(workspace);
    float* means = reinterpret_cast<float*>(ws + (size_t)B * plan.L.stride);
    const int ucap = (N + 3) / 4 * 4;
    const size_t smem = (size_t)plan.P * (sizeof(unsigned long long) + sizeof(int)) + (size_t)(3 * N + 4) * sizeof(int) +
                        (size_t)kMeanChans * (N + 4 + ucap) * sizeof(float);
    auto kern = shape == 2 ? vox_front_kernel<true> : vox_front_kernel<false>;
    RI_KERNEL_SETUP(kern, true, ri_step_carveout_percent());
    dim3 grid(C > 0 ? (C + kMeanChans - 1) / kMeanChans : 1, B);
    // norm_mode bit 8: compute the per-cloud mean inside the kernel (torch's reduction order) and write it to `mean`
    // norm_mode bit 9: `edge` is [B,C,N] and receives only the feat - mean(cell) half of the edge features
    const int own_mean = (norm_mode & 0x100) != 0;
    norm_mode &= 0x2ff;
    if (own_mean && (N < 128 || (N & 3) != 0 || ((uintptr_t)points & 15) != 0)) return RI_ERR_UNSUPPORTED;
    const float mean_factor = (float)(3LL * B) / (float)(3LL * B * N);           // mean_kernel_cuda: num_outputs / numel
    kern<<<grid, kFrontThreads, smem, (cudaStream_t)stream>>>(points, pstride, const_cast<float*>(mean), own_mean, mean_factor,
                                                              feat, C, N, plan.P, r, (int)s_ll,
                                                              plan.tile_cells, plan.ntiles, shape, eps, norm_mode, ucap,
                                                              norm_coords, vox_coords, ind, ws, means, edge);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
