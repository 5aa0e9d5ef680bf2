// gridsub.cu — barycentre grid subsampling of a scan on the GPU, sm_100a  (SURVEY.md §8 row f2).
//
// Replaces  grid_subsampling()  (/root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:4-106,
//           accumulator class grid_subsampling.h:9-84, PointXYZ arithmetic cpp_wrappers/cpp_utils/cloud/cloud.h:40-155)
// which the reference runs on the host through a CPython extension (cpp_subsampling/wrapper.cpp:58-285,
// utils/grid_subsampleing.py:3-21), one point after another into an unordered_map.
//
// Semantics kept (all fp32 unless noted):
//   origin      = floor(min_corner * (1.0f / dl)) * dl                                   (:25-27)
//   nx, ny      = (size_t)floor((max_corner - origin) / dl) + 1                          (:30-31)
//   cell(p)     = ix + nx*iy + nx*ny*iz,  i* = (size_t)floor((p.* - origin.*) / dl)      (:53-56)
//   barycentre  = (sum of the cell's points, added in ORIGINAL point order) * (float)(1.0 / count)   (:84)
//   features    = (sum in original order) / (float)count                                  (:87-92)
//   labels      = per label column the most frequent value of the cell                    (:96-100)
// The reference emits cells in unordered_map iteration order (unspecified); this implementation emits them in
// ASCENDING cell index, and breaks label-count ties towards the SMALLEST label (the reference's tie-break is the
// map's iteration order, i.e. unspecified as well).  Everything else is bit-identical: the cell's points are summed
// in original order because the sort below is stable.
//
// Pipeline (one stream, no host synchronisation, output count left in device memory):
//   bounds   min / max corner: block reduction + atomicMin/Max on order-preserving integer images of the floats
//   keys     64-bit cell index per point (+ iota), geometry recomputed per thread from the 6 bounds; the largest key of
//            the scan is kept in device memory (atomicMax)
//   sort     this repo's stable LSD radix sort (csrc/radix.cuh) in the role of the reference's unordered_map: 8 passes are
//            enqueued for the 64-bit keys, the passes above the largest key's top bit return at once (a 400k-point scan
//            on a 4.5 cm grid has ~21 key bits: 3 passes run)
//   heads    heads of the cell runs counted per block of 1024 sorted points, then every block sums the counts of the
//            blocks before it and scans its own head flags: output slot of every cell; slot count = M
//   reduce   one thread per (cell, channel) walks the cell's points in sorted (= original) order
#include "ri_common.cuh"
#include "radix.cuh"

namespace {

constexpr int kGsThreads = 256;

__device__ __forceinline__ unsigned gs_ord(float f)          // monotone float -> uint
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float gs_unord(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct GsGeom { float ox, oy, oz; unsigned long long nx, ny; };

__device__ __forceinline__ GsGeom gs_geom(const unsigned* __restrict__ bounds, float dl)
{
    GsGeom g;
    const float inv = __fdiv_rn(1.0f, dl);                                           // (1/sampleDl): int / float -> float
    const float mnx = gs_unord(bounds[0]), mny = gs_unord(bounds[1]), mnz = gs_unord(bounds[2]);
    const float mxx = gs_unord(bounds[3]), mxy = gs_unord(bounds[4]);
    g.ox = __fmul_rn(floorf(__fmul_rn(mnx, inv)), dl);
    g.oy = __fmul_rn(floorf(__fmul_rn(mny, inv)), dl);
    g.oz = __fmul_rn(floorf(__fmul_rn(mnz, inv)), dl);
    g.nx = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxx, g.ox), dl)) + 1ull;
    g.ny = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxy, g.oy), dl)) + 1ull;
    return g;
}

__global__ void gs_init_kernel(unsigned* bounds, int* out_count)
{
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
    if (threadIdx.x == 6) *out_count = 0;
    if (threadIdx.x == 7) *reinterpret_cast<unsigned long long*>(bounds + 8) = 0ull;          // the largest key
}

__global__ void __launch_bounds__(kGsThreads)
gs_bounds_kernel(const float* __restrict__ pts, int N, unsigned* __restrict__ bounds)
{
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (size_t i = (size_t)blockIdx.x * kGsThreads + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * kGsThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned o = gs_ord(pts[3 * i + a]);
            lo[a] = min(lo[a], o); hi[a] = max(hi[a], o);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(bounds + a, lo[a]); atomicMax(bounds + 3 + a, hi[a]); }
    }
}

__global__ void __launch_bounds__(kGsThreads)
gs_keys_kernel(const float* __restrict__ pts, int N, float dl, const unsigned* __restrict__ bounds,
               unsigned long long* __restrict__ keys, int* __restrict__ vals, unsigned long long* __restrict__ maxkey)
{
    const int i = blockIdx.x * kGsThreads + threadIdx.x;
    unsigned long long key = 0ull;
    if (i < N) {
        const GsGeom g = gs_geom(bounds, dl);
        const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 0], g.ox), dl));
        const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 1], g.oy), dl));
        const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 2], g.oz), dl));
        key = ix + g.nx * iy + g.nx * g.ny * iz;
        keys[i] = key;
        vals[i] = i;
    }
    // the largest key decides how many radix passes run (radix.cuh): one atomic per warp
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    lo = hi == mhi ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
    if ((threadIdx.x & 31) == 0) atomicMax(maxkey, ((unsigned long long)mhi << 32) | mlo);
}

// Heads of the cell runs (a sorted point whose key differs from its predecessor's), kGsHeadItems points per thread.
constexpr int kGsHeadItems = 4;
constexpr int kGsHeadBlock = kGsThreads * kGsHeadItems;

__device__ __forceinline__ const unsigned long long* gs_sorted_keys(const unsigned long long* ka, const unsigned long long* kb,
                                                                   const unsigned long long* maxkey)
{
    return ri_radix::rs_result_in_b(maxkey, 64) ? kb : ka;
}

__global__ void __launch_bounds__(kGsThreads)
gs_count_kernel(const unsigned long long* __restrict__ ka, const unsigned long long* __restrict__ kb,
                const unsigned long long* __restrict__ maxkey, int N, int* __restrict__ blockcnt)
{
    __shared__ int swarp[kGsThreads / 32];
    const unsigned long long* keys = gs_sorted_keys(ka, kb, maxkey);
    const int base = blockIdx.x * kGsHeadBlock + threadIdx.x * kGsHeadItems;
    int c = 0;
#pragma unroll
    for (int e = 0; e < kGsHeadItems; ++e) {
        const int i = base + e;
        if (i < N && (i == 0 || keys[i] != keys[i - 1])) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kGsThreads / 32; ++w) t += swarp[w];
        blockcnt[blockIdx.x] = t;
    }
}

// starts[slot] = first sorted position of output cell `slot`; starts[M] = N; *out_count = M
__global__ void __launch_bounds__(kGsThreads)
gs_starts_kernel(const unsigned long long* __restrict__ ka, const unsigned long long* __restrict__ kb,
                 const unsigned long long* __restrict__ maxkey, int N, const int* __restrict__ blockcnt, int nblocks,
                 int* __restrict__ starts, int* __restrict__ out_count)
{
    __shared__ int swarp[kGsThreads / 32];
    __shared__ int sbefore;
    const unsigned long long* keys = gs_sorted_keys(ka, kb, maxkey);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // heads in the blocks before this one
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kGsThreads) part += blockcnt[b];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) swarp[w] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int q = 0; q < kGsThreads / 32; ++q) t += swarp[q];
        sbefore = t;
    }
    __syncthreads();
    const int before = sbefore;
    __syncthreads();
    // this block's head flags, scanned in position order
    const int base = blockIdx.x * kGsHeadBlock + threadIdx.x * kGsHeadItems;
    bool head[kGsHeadItems];
    int c = 0;
#pragma unroll
    for (int e = 0; e < kGsHeadItems; ++e) {
        const int i = base + e;
        head[e] = i < N && (i == 0 || keys[i] != keys[i - 1]);
        c += head[e] ? 1 : 0;
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) swarp[w] = incl;
    __syncthreads();
    int woff = 0;
    for (int q = 0; q < w; ++q) woff += swarp[q];
    int slot = before + woff + incl - c;
#pragma unroll
    for (int e = 0; e < kGsHeadItems; ++e)
        if (head[e]) starts[slot++] = base + e;
    if (blockIdx.x == (unsigned)(nblocks - 1) && threadIdx.x == kGsThreads - 1) { *out_count = slot; starts[slot] = N; }
}

// one thread per (cell, channel): channel 0..2 = barycentre, 3..3+fdim-1 = features, then ldim label columns
__global__ void __launch_bounds__(kGsThreads)
gs_reduce_kernel(const float* __restrict__ pts, const float* __restrict__ feats, const int* __restrict__ labels,
                 const int* __restrict__ order_a, const int* __restrict__ order_b, const unsigned long long* __restrict__ maxkey,
                 const int* __restrict__ starts, const int* __restrict__ count_ptr,
                 int fdim, int ldim, float* __restrict__ out_pts, float* __restrict__ out_feats, int* __restrict__ out_labels)
{
    const int* order = ri_radix::rs_result_in_b(maxkey, 64) ? order_b : order_a;
    const int M = *count_ptr;
    const int chans = 3 + fdim + ldim;
    const long long total = (long long)M * chans;
    for (long long w = (long long)blockIdx.x * kGsThreads + threadIdx.x; w < total; w += (long long)gridDim.x * kGsThreads) {
        const int cell = (int)(w / chans), ch = (int)(w - (long long)cell * chans);
        const int s0 = starts[cell], s1 = starts[cell + 1];
        const int cnt = s1 - s0;
        if (ch < 3) {
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc = __fadd_rn(acc, pts[3 * (size_t)order[s] + ch]);
            out_pts[3 * (size_t)cell + ch] = __fmul_rn(acc, (float)(1.0 / (double)cnt));        // point * (1.0 / count)
        } else if (ch < 3 + fdim) {
            const int f = ch - 3;
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc = __fadd_rn(acc, feats[(size_t)order[s] * fdim + f]);
            out_feats[(size_t)cell * fdim + f] = __fdiv_rn(acc, (float)cnt);                    // f / (float)count
        } else {
            // most frequent label of the cell, the smallest one among equally frequent ones.  Small cells: count every label
            // against the cell (cnt^2 <= 4096 compares).  Large cells (a coarse grid on a 50k-point scan): one pass that
            // counts into a 64-slot open-addressing table in local memory — O(cnt) while the cell holds at most 48 distinct
            // labels (semantic classes); a cell with more falls back to the quadratic count.
            const int l = ch - 3 - fdim;
            int best = 0, best_n = 0;
            bool done = false;
            if (cnt > 64) {
                int hk[64], hc[64];
                for (int q = 0; q < 64; ++q) hc[q] = 0;
                int distinct = 0;
                bool overflow = false;
                for (int s = s0; s < s1 && !overflow; ++s) {
                    const int v = labels[(size_t)order[s] * ldim + l];
                    unsigned h = ((unsigned)v * 2654435761u) >> 26;
                    while (hc[h] != 0 && hk[h] != v) h = (h + 1) & 63;
                    if (hc[h] == 0) { if (++distinct > 48) { overflow = true; break; } hk[h] = v; }
                    ++hc[h];
                }
                if (!overflow) {
                    for (int q = 0; q < 64; ++q)
                        if (hc[q] > best_n || (hc[q] == best_n && hc[q] > 0 && hk[q] < best)) { best = hk[q]; best_n = hc[q]; }
                    done = true;
                }
            }
            if (!done) {
                best = 0; best_n = 0;
                for (int s = s0; s < s1; ++s) {
                    const int v = labels[(size_t)order[s] * ldim + l];
                    int n = 0;
                    for (int t = s0; t < s1; ++t) n += labels[(size_t)order[t] * ldim + l] == v ? 1 : 0;
                    if (n > best_n || (n == best_n && v < best)) { best = v; best_n = n; }
                }
            }
            out_labels[(size_t)cell * ldim + l] = best;
        }
    }
}

struct GsLayout {
    size_t bounds, keys_in, keys_out, vals_in, vals_out, blockcnt, starts, hist, total;
};

static size_t gs_align(size_t x) { return (x + 255) & ~(size_t)255; }

static int gs_layout(int N, GsLayout& L)
{
    size_t off = 0;
    L.bounds = off; off += gs_align(8 * sizeof(unsigned) + sizeof(unsigned long long));      // 6 bounds, pad, the largest key
    L.keys_in = off; off += gs_align((size_t)N * 8);
    L.keys_out = off; off += gs_align((size_t)N * 8);
    L.vals_in = off; off += gs_align((size_t)N * 4);
    L.vals_out = off; off += gs_align((size_t)N * 4);
    L.blockcnt = off; off += gs_align(((size_t)N / kGsHeadBlock + 1) * 4);
    L.starts = off; off += gs_align(((size_t)N + 1) * 4);
    L.hist = off; off += gs_align(ri_radix::hist_bytes(N));
    L.total = off;
    return RI_OK;
}

}  // namespace

extern "C" size_t ri_grid_subsample_workspace_bytes(int N)
{
    if (N <= 0) return 256;
    GsLayout L;
    if (gs_layout(N, L) != RI_OK) return 0;
    return L.total;
}

// points [N,3], features [N,fdim] or NULL (fdim = 0), labels [N,ldim] or NULL (ldim = 0), all row-major as the reference's numpy
// arrays; out_* sized for N cells (the worst case); *out_count (device memory) receives the number of cells M.
extern "C" int ri_grid_subsample_f32(const float* points, const float* features, const int* labels, int N, int fdim, int ldim,
                                     float dl, float* out_points, float* out_features, int* out_labels, int* out_count,
                                     void* workspace, size_t workspace_bytes, void* stream)
{
    if (N < 0 || fdim < 0 || ldim < 0 || !(dl > 0.0f) || out_count == nullptr) return RI_ERR_BAD_ARG;
    if ((fdim > 0 && (features == nullptr || out_features == nullptr)) || (ldim > 0 && (labels == nullptr || out_labels == nullptr)))
        return RI_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { cudaError_t e = cudaMemsetAsync(out_count, 0, sizeof(int), st); return e == cudaSuccess ? RI_OK : (int)e; }
    GsLayout L;
    int rc = gs_layout(N, L);
    if (rc != RI_OK) return rc;
    if (workspace == nullptr || workspace_bytes < L.total) return RI_ERR_WORKSPACE;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    unsigned* bounds = reinterpret_cast<unsigned*>(ws + L.bounds);
    unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(ws + L.keys_in);
    unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws + L.keys_out);
    int* vals_in = reinterpret_cast<int*>(ws + L.vals_in);
    int* vals_out = reinterpret_cast<int*>(ws + L.vals_out);
    int* blockcnt = reinterpret_cast<int*>(ws + L.blockcnt);
    int* starts = reinterpret_cast<int*>(ws + L.starts);
    int* hist = reinterpret_cast<int*>(ws + L.hist);
    unsigned long long* maxkey = reinterpret_cast<unsigned long long*>(bounds + 8);
    const int blocks = (N + kGsThreads - 1) / kGsThreads;
    const int hblocks = (N + kGsHeadBlock - 1) / kGsHeadBlock;
    const int sms = ri_num_sms();

    gs_init_kernel<<<1, 32, 0, st>>>(bounds, out_count);
    gs_bounds_kernel<<<min(blocks, 8 * sms), kGsThreads, 0, st>>>(points, N, bounds);
    gs_keys_kernel<<<blocks, kGsThreads, 0, st>>>(points, N, dl, bounds, keys_in, vals_in, maxkey);
    RI_LAUNCH_CHECK();
    const int rc_sort = ri_radix::sort_pairs<unsigned long long>(keys_in, keys_out, vals_in, vals_out, N, 64, maxkey, hist, st);
    if (rc_sort != RI_OK) return rc_sort;
    gs_count_kernel<<<hblocks, kGsThreads, 0, st>>>(keys_in, keys_out, maxkey, N, blockcnt);
    gs_starts_kernel<<<hblocks, kGsThreads, 0, st>>>(keys_in, keys_out, maxkey, N, blockcnt, hblocks, starts, out_count);
    const long long work = (long long)N * (3 + fdim + ldim);
    const int rblocks = (int)min((long long)32 * sms, (work + kGsThreads - 1) / kGsThreads);
    gs_reduce_kernel<<<rblocks, kGsThreads, 0, st>>>(points, features, labels, vals_in, vals_out, maxkey, starts, out_count, fdim,
                                                      ldim, out_points, out_features, out_labels);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
