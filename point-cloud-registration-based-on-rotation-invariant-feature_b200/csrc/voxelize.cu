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

namespace {

constexpr int kPrepThreads = 512;
constexpr int kSmallCloudMax = 4096;     // K1 sorts a whole cloud inside one CTA
constexpr int kTileCells = 8192;         // 32 KB of fp32 per tile
constexpr int kRing = 3;                 // tiles in flight per CTA
constexpr int kFillThreads = 256;

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
            if (cell >= 0 && cell < s) key = ((unsigned long long)(unsigned)cell << 32) | (unsigned)i;
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
            if (key != ~0ull) {
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
            if (key != ~0ull) {
                W[L.off_pid + u] = (int)(unsigned)(key & 0xffffffffu);
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
__global__ void __launch_bounds__(kFillThreads, 2)
vox_fill_kernel(const float* __restrict__ feat, const int* __restrict__ ws, int B, int C, int N, int s,
                int tile_cells, int ntiles, int planes_per_item, int ngroups,
                float* __restrict__ out, int* __restrict__ cnt)
{
    extern __shared__ __align__(128) float sring[];        // kRing tiles of tile_cells floats
    const int tid = threadIdx.x;
    const VoxWs L = vox_ws_layout(N, ntiles);
    const int planes = C + 1;                              // plane C is the integer count grid

    for (int i = tid; i < kRing * tile_cells / 4; i += kFillThreads)
        reinterpret_cast<float4*>(sring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const long long total = (long long)B * ntiles * ngroups;
    int slot = 0;
    int prev_sA = 0, prev_sB = 0, prev_lo = 0;
    const int* prevW = nullptr;

    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        // item -> (cloud, tile, plane group); groups of one (cloud,tile) are adjacent so neighbouring CTAs
        // share the occupied-cell table in L2.
        const int g = (int)(item % ngroups);
        const long long bt = item / ngroups;
        const int t = (int)(bt % ntiles);
        const int b = (int)(bt / ntiles);
        const int* W = ws + (size_t)b * L.stride;
        const int cell_lo = t * tile_cells;
        const int ncell = min(tile_cells, s - cell_lo);
        const int sA = __ldg(W + L.off_tile + t), sB = __ldg(W + L.off_tile + t + 1);
        const int p0 = g * planes_per_item, p1 = min(planes, p0 + planes_per_item);

        // un-patch the ring: the previous item's occupied cells are the only non-zero words in it
        if (tid == 0) ri_bulk_wait_read<0>();
        __syncthreads();
        if (prevW != nullptr) {
            for (int sg = prev_sA + tid; sg < prev_sB; sg += kFillThreads) {
                const int off = __ldg(prevW + L.off_cell + sg) - prev_lo;
#pragma unroll
                for (int q = 0; q < kRing; ++q) sring[q * tile_cells + off] = 0.f;
            }
        }
        prevW = W; prev_sA = sA; prev_sB = sB; prev_lo = cell_lo;
        // (the __syncthreads at the top of the plane loop orders these stores before the patching)

        for (int p = p0; p < p1; ++p) {
            float* tile = sring + slot * tile_cells;
            if (tid == 0) ri_bulk_wait_read<kRing - 1>();  // the copy that last used this slot has left smem
            __syncthreads();
            const float* F = feat + ((size_t)b * C + (p < C ? p : 0)) * N;
            for (int sg = sA + tid; sg < sB; sg += kFillThreads) {
                const int cell = __ldg(W + L.off_cell + sg);
                const int st = __ldg(W + L.off_start + sg), en = __ldg(W + L.off_start + sg + 1);
                float val;
                if (p < C) {
                    const float inv = __fdiv_rn(1.0f, (float)(en - st));               // vox.cu:66
                    float acc = 0.f;
                    for (int u = st; u < en; ++u)
                        acc = __fadd_rn(acc, __fmul_rn(__ldg(F + __ldg(W + L.off_pid + u)), inv));
                    val = acc;
                } else {
                    val = __int_as_float(en - st);
                }
                tile[cell - cell_lo] = val;
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
                  float* out, int* ind, int* cnt, void* workspace, size_t ws_bytes, cudaStream_t st)
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
        const int planes = C + 1;
        const long long bt = (long long)B * plan.ntiles;
        long long grid = 2LL * sms;
        long long want_groups = (8 * grid + bt - 1) / bt;          // ~8 items per CTA for balance
        if (want_groups < 1) want_groups = 1;
        if (want_groups > planes) want_groups = planes;
        int ppi = (int)((planes + want_groups - 1) / want_groups);
        if (ppi < 2 && planes >= 2) ppi = 2;
        const int ngroups = (planes + ppi - 1) / ppi;
        const long long total = bt * ngroups;
        if (grid > total) grid = total;
        vox_fill_kernel<<<(unsigned)grid, kFillThreads, smem2, st>>>(feat, ws, B, C, N, s, plan.tile_cells, plan.ntiles,
                                                                     ppi, ngroups, out, cnt);
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
    return voxelize_impl<true>(feat, coords, B, C, N, r, out, ind, cnt, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ri_cube_voxelize_f32(const float* feat, const int* coords, int B, int C, int N, int r,
                                    float* out, int* ind, int* cnt, void* workspace, size_t ws_bytes, void* stream)
{
    return voxelize_impl<false>(feat, coords, B, C, N, r, out, ind, cnt, workspace, ws_bytes, (cudaStream_t)stream);
}
