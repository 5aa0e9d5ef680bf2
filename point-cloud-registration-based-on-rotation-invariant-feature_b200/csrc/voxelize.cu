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
//   K2  vox_fill      persistent CTAs (2 per SM).  A work item is (cloud, grid tile, group of channel planes).
//                     Each CTA keeps a ring of zeroed 32 KB tiles in shared memory; per plane it patches the
//                     few occupied cells of the tile with their means (one thread per cell, sequential sum),
//                     and ships the tile to HBM with ONE bulk async copy (cp.async.bulk shared->global, the
//                     TMA engine, SASS UBLKCP), overlapping the next plane's patching with the store.  The
//                     integer count grid goes out the same way as plane C.  Nothing is memset, nothing is
//                     read back, `out` is written exactly once, and there are no atomics at all.
//
//   Fallback (N > 4096 points per cloud, r^3 not a multiple of 4, or misaligned pointers): memset + integer
//   atomics for counts + float atomics for the means over a (point tile, channel group, cloud) grid.
#include "ri_common.cuh"
#include <stdlib.h>

extern "C" int ri_voxel_edge_gather_f32(const float* avg, const float* feat, const int* inds, int B, int C, int N, int s,
                                        float* out, void* stream);

namespace {

constexpr int kPrepThreads = 512;
constexpr int kSmallCloudMax = 4096;     // K1 sorts a whole cloud inside one CTA
constexpr int kTileCells = 8192;         // 32 KB of fp32 per tile
constexpr int kRing = 3;                 // tiles in flight per CTA
constexpr int kFillThreads = 256;
constexpr int kPlaneBatch = 8;            // channel planes whose gathers are issued together
constexpr int kFillMaxRegs = 96;          // 2 CTAs/SM use 48K registers: leaves room for the k-NN CTAs of the other branch
constexpr int kSegCache = 2;             // occupied cells per thread whose table entries live in registers
constexpr unsigned kNoCell = 0xffffffffu;

struct VoxWs {                            // per-cloud int32 workspace layout
    int stride, off_pid, off_cell, off_start, off_tile, off_meta;
};
__host__ __device__ inline VoxWs vox_ws_layout(int N, int ntiles)
{
    VoxWs w;
    w.off_pid = 0;
    w.off_cell = N;
    w.off_start = 2 * N;
    w.off_tile = 3 * N + 1;
    w.off_meta = w.off_tile + ntiles + 1;
    w.stride = (w.off_meta + 2 + 3) / 4 * 4;
    return w;
}

// ------------------------------------------------------------------------------------------------ K1
// coords: float [B,3,N] (SPH) or int [B,3,N] (cube).  P = power of two >= N.
template <bool SPH>
__global__ void __launch_bounds__(kPrepThreads)
vox_prepare_kernel(const void* __restrict__ coords_v, int N, int P, int r, int s, int tile_cells, int ntiles,
                   int* __restrict__ ind, int* __restrict__ ws)
{
    extern __shared__ unsigned long long skeys[];          // [P] keys, then [P] ints of segment cells
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
// One occupied cell ("segment" of the cell-sorted point list) of the current tile, as cached by its thread.
struct SegRegs { int off, st, cnt; };

// Mean of channel plane p over one segment (ascending point order), or the count itself for plane C.
// When `edge` is given also emits the DGCNN edge features of the segment's points:
//   edge[b, p, i] = feat - mean,  edge[b, C+p, i] = feat      (pvconv.py:68-90; undefined points: see below)
__device__ __forceinline__ float seg_value(const SegRegs sg, int p, int C, int N, const float* __restrict__ Fp,
                                           const int* __restrict__ pid, float* __restrict__ edge_rel,
                                           float* __restrict__ edge_cpy)
{
    if (p >= C) return __int_as_float(sg.cnt);
    const float inv = __fdiv_rn(1.0f, (float)sg.cnt);                               // vox.cu:66
    float acc = 0.f;
    for (int u = sg.st; u < sg.st + sg.cnt; ++u)
        acc = __fadd_rn(acc, __fmul_rn(__ldg(Fp + __ldg(pid + u)), inv));          // vox.cu:68-70, point order
    if (edge_rel != nullptr) {
        for (int u = sg.st; u < sg.st + sg.cnt; ++u) {
            const int i = __ldg(pid + u);
            const float f = __ldg(Fp + i);
            edge_rel[i] = __fsub_rn(f, acc);
            edge_cpy[i] = f;
        }
    }
    return acc;
}

__global__ void __maxnreg__(kFillMaxRegs)
vox_fill_kernel(const float* __restrict__ feat, const int* __restrict__ ws, int B, int C, int N, int s,
                int tile_cells, int ntiles,
                float* __restrict__ out, int* __restrict__ cnt, float* __restrict__ edge)
{
    extern __shared__ __align__(128) float sring[];        // kRing tiles of tile_cells floats
    const int tid = threadIdx.x;
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int planes = C + 1;                              // plane C is the integer count grid

    for (int i = tid; i < kRing * tile_cells / 4; i += kFillThreads)
        reinterpret_cast<float4*>(sring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // Work = the flattened sequence of (cloud, tile, plane) units, 32 KB of output each.  Every CTA takes one
    // contiguous, equally long slice of it: perfectly balanced, and a CTA changes (cloud, tile) at most
    // ceil(slice / planes) + 1 times, so the per-item table fetch is paid once or twice per CTA, not per group.
    const long long total = (long long)B * ntiles * planes;
    const long long u_begin = total * blockIdx.x / gridDim.x;
    const long long u_end = total * (blockIdx.x + 1) / gridDim.x;
    int slot = 0;
    // The ring slots are never re-zeroed wholesale.  Between two bulk copies out of a slot only the occupied cells
    // of the tile are rewritten; when the CTA moves on to another (cloud, tile) the cells of the PREVIOUS item are
    // cleared lazily, slot by slot, right before each slot's first reuse — so the copies still in flight are never
    // waited for (no pipeline drain at an item switch).
    SegRegs seg[kSegCache];                                 // this item's cells owned by this thread
    int old_off[kSegCache];                                 // previous item's cells (offset, or -1)
#pragma unroll
    for (int q = 0; q < kSegCache; ++q) { seg[q].off = 0; seg[q].st = 0; seg[q].cnt = 0; old_off[q] = -1; }
    int extra_lo = 0, extra_hi = 0, cur_cell_lo = 0;        // cells beyond the register cache (rare), current item
    int old_extra_lo = 0, old_extra_hi = 0, old_cell_lo = 0;
    const int* curW = nullptr;
    const int* oldW = nullptr;
    int stale = 0;                                          // ring slots that still carry the previous item's cells
    int prev_planes = kRing;

    auto unpatch_slot = [&](float* tile) {
#pragma unroll
        for (int q = 0; q < kSegCache; ++q)
            if (old_off[q] >= 0) tile[old_off[q]] = 0.f;
        for (int sg = old_extra_lo + tid; sg < old_extra_hi; sg += kFillThreads)
            tile[__ldg(oldW + L.off_cell + sg) - old_cell_lo] = 0.f;
    };

    for (long long u = u_begin; u < u_end;) {
        const long long bt = u / planes;
        const int p0 = (int)(u - bt * planes);
        const int p1 = (int)min((long long)planes, p0 + (u_end - u));
        u += p1 - p0;
        const int t = (int)(bt % ntiles);
        const int b = (int)(bt / ntiles);
        const int* W = ws + (size_t)b * L.stride;
        const int* pid = W + L.off_pid;
        const int cell_lo = t * tile_cells;
        const int ncell = min(tile_cells, s - cell_lo);
        const int sA = __ldg(W + L.off_tile + t), sB = __ldg(W + L.off_tile + t + 1);

        if (curW != nullptr) {
            if (stale > 0 || prev_planes < kRing) {
                // the previous item did not cycle through the whole ring: clean everything the slow way
                if (tid == 0) ri_bulk_wait_read<0>();
                __syncthreads();
                for (int k2 = 0; k2 < kRing; ++k2) {
                    float* tile = sring + k2 * tile_cells;
                    if (stale > 0) unpatch_slot(tile);
#pragma unroll
                    for (int q = 0; q < kSegCache; ++q)
                        if (seg[q].cnt > 0) tile[seg[q].off] = 0.f;
                    for (int sg = extra_lo + tid; sg < extra_hi; sg += kFillThreads)
                        tile[__ldg(curW + L.off_cell + sg) - cur_cell_lo] = 0.f;
                }
                __syncthreads();
                stale = 0;
#pragma unroll
                for (int q = 0; q < kSegCache; ++q) old_off[q] = -1;
                old_extra_lo = old_extra_hi = 0;
            } else {
#pragma unroll
                for (int q = 0; q < kSegCache; ++q) old_off[q] = seg[q].cnt > 0 ? seg[q].off : -1;
                old_extra_lo = extra_lo; old_extra_hi = extra_hi; old_cell_lo = cur_cell_lo; oldW = curW;
                stale = kRing;
            }
        }
        // this item's table entries -> registers
#pragma unroll
        for (int q = 0; q < kSegCache; ++q) {
            const int sg = sA + tid + q * kFillThreads;
            seg[q].cnt = 0;
            if (sg < sB) {
                seg[q].off = __ldg(W + L.off_cell + sg) - cell_lo;
                seg[q].st = __ldg(W + L.off_start + sg);
                seg[q].cnt = __ldg(W + L.off_start + sg + 1) - seg[q].st;
            }
        }
        extra_lo = min(sB, sA + kSegCache * kFillThreads); extra_hi = sB; cur_cell_lo = cell_lo; curW = W;
        prev_planes = p1 - p0;

        // undefined points of this cloud (ind == -1): edge rows are (0, feat); done by whoever owns tile 0's planes
        if (edge != nullptr && t == 0) {
            const int nvalid = __ldg(W + L.off_meta + 1);
            for (int u = nvalid + tid; u < N; u += kFillThreads) {
                const int i = __ldg(pid + u);
                for (int p = p0; p < min(p1, C); ++p) {
                    edge[((size_t)b * 2 * C + p) * N + i] = 0.f;
                    edge[((size_t)b * 2 * C + C + p) * N + i] = __ldg(feat + ((size_t)b * C + p) * N + i);
                }
            }
        }

        // Planes are handled in batches of kPlaneBatch: the gathers of a whole batch (independent loads, one per
        // plane and point) are issued together, so their latency is paid once per batch, not once per 32 KB tile.
        for (int pb = p0; pb < p1; pb += kPlaneBatch) {
            float val[kSegCache][kPlaneBatch];
#pragma unroll
            for (int q = 0; q < kSegCache; ++q) {
                if (seg[q].cnt <= 0) continue;
                const float inv = __fdiv_rn(1.0f, (float)seg[q].cnt);                       // vox.cu:66
#pragma unroll
                for (int j = 0; j < kPlaneBatch; ++j) val[q][j] = 0.f;
                for (int u = seg[q].st; u < seg[q].st + seg[q].cnt; ++u) {                  // ascending point order
                    const int i = __ldg(pid + u);
                    float f[kPlaneBatch];
#pragma unroll
                    for (int j = 0; j < kPlaneBatch; ++j)
                        f[j] = (pb + j < min(p1, C)) ? __ldg(feat + ((size_t)b * C + pb + j) * N + i) : 0.f;
#pragma unroll
                    for (int j = 0; j < kPlaneBatch; ++j)
                        val[q][j] = __fadd_rn(val[q][j], __fmul_rn(f[j], inv));             // vox.cu:68-70
                }
                if (edge != nullptr) {
                    for (int u = seg[q].st; u < seg[q].st + seg[q].cnt; ++u) {
                        const int i = __ldg(pid + u);
#pragma unroll
                        for (int j = 0; j < kPlaneBatch; ++j)
                            if (pb + j < min(p1, C)) {
                                const float f = __ldg(feat + ((size_t)b * C + pb + j) * N + i);
                                edge[((size_t)b * 2 * C + pb + j) * N + i] = __fsub_rn(f, val[q][j]);
                                edge[((size_t)b * 2 * C + C + pb + j) * N + i] = f;
                            }
                    }
                }
#pragma unroll
                for (int j = 0; j < kPlaneBatch; ++j)
                    if (pb + j >= C) val[q][j] = __int_as_float(seg[q].cnt);               // the count plane
            }
#pragma unroll
            for (int j = 0; j < kPlaneBatch; ++j) {
                const int p = pb + j;
                if (p >= p1) break;
                float* tile = sring + slot * tile_cells;
                if (tid == 0) ri_bulk_wait_read<kRing - 1>();  // the copy that last used this slot has left smem
                __syncthreads();
                if (stale > 0) {                                // first reuse of this slot since the item switch
                    unpatch_slot(tile);
                    --stale;
                    __syncthreads();                            // an old cell may coincide with a new one of another thread
                }
#pragma unroll
                for (int q = 0; q < kSegCache; ++q)
                    if (seg[q].cnt > 0) tile[seg[q].off] = val[q][j];
                if (extra_lo < extra_hi) {                      // cells beyond the register cache: direct path
                    const float* Fp = feat + ((size_t)b * C + (p < C ? p : 0)) * N;
                    float* er = (edge != nullptr && p < C) ? edge + ((size_t)b * 2 * C + p) * N : nullptr;
                    float* ec = (edge != nullptr && p < C) ? edge + ((size_t)b * 2 * C + C + p) * N : nullptr;
                    for (int sg = extra_lo + tid; sg < extra_hi; sg += kFillThreads) {
                        SegRegs x;
                        x.off = __ldg(W + L.off_cell + sg) - cell_lo;
                        x.st = __ldg(W + L.off_start + sg);
                        x.cnt = __ldg(W + L.off_start + sg + 1) - x.st;
                        tile[x.off] = seg_value(x, p, C, N, Fp, pid, er, ec);
                    }
                }
                ri_fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    void* dst = (p < C) ? (void*)(out + ((size_t)b * C + p) * s + cell_lo)
                                        : (void*)(cnt + (size_t)b * s + cell_lo);
                    ri_bulk_store(dst, tile, (uint32_t)ncell * 4u);
                    ri_bulk_commit();
                }
                slot = (slot + 1 == kRing) ? 0 : slot + 1;
            }
        }
    }
    if (tid == 0) ri_bulk_wait<0>();
}

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

VoxPlan vox_plan(int N, int r, const void* out, const void* cnt)
{
    VoxPlan p;
    const long long s = (long long)r * r * r;
    p.tiled = N <= kSmallCloudMax && (s % 4 == 0) &&
              ((uintptr_t)out % 16 == 0) && ((uintptr_t)cnt % 16 == 0);
    p.tile_cells = (int)(s < kTileCells ? s : kTileCells);
    p.ntiles = (int)((s + p.tile_cells - 1) / p.tile_cells);
    p.P = next_pow2(N < 2 ? 2 : N);
    p.L = vox_ws_layout(N, p.ntiles);
    return p;
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
        const size_t need = (size_t)B * plan.L.stride * sizeof(int);
        if (workspace == nullptr || ws_bytes < need) return RI_ERR_WORKSPACE;
        int* ws = reinterpret_cast<int*>(workspace);
        const size_t smem1 = (size_t)plan.P * (sizeof(unsigned long long) + sizeof(int));
        cudaError_t e;
        if (smem1 + 1024 > 48 * 1024) {
            e = cudaFuncSetAttribute(vox_prepare_kernel<SPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
            if (e != cudaSuccess) return (int)e;
        }
        vox_prepare_kernel<SPH><<<B, kPrepThreads, smem1, st>>>(coords, N, plan.P, r, s, plan.tile_cells, plan.ntiles, ind, ws);
        RI_LAUNCH_CHECK();

        const int sms = ri_num_sms();
        const size_t smem2 = (size_t)kRing * plan.tile_cells * sizeof(float);
        e = cudaFuncSetAttribute(vox_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        const long long units = (long long)B * plan.ntiles * (C + 1);
        long long grid = 2LL * sms;
        if (grid > units) grid = units;
        vox_fill_kernel<<<(unsigned)grid, kFillThreads, smem2, st>>>(feat, ws, B, C, N, s, plan.tile_cells, plan.ntiles,
                                                                     out, cnt, edge);
        RI_LAUNCH_CHECK();
        return RI_OK;
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

extern "C" size_t ri_voxelize_workspace_bytes(int B, int N, int r)
{
    if (B <= 0 || N <= 0 || r <= 0) return 16;
    const long long s = (long long)r * r * r;
    const int tile_cells = (int)(s < kTileCells ? s : kTileCells);
    const int ntiles = (int)((s + tile_cells - 1) / tile_cells);
    return (size_t)B * vox_ws_layout(N, ntiles).stride * sizeof(int) + 16;
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
