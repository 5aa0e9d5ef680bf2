// knn_grid.cu — k-nearest neighbours through a uniform hash grid, for scan-sized clouds (~50k points), sm_100a.
//
// Same contract as knn.cu / the reference's KnnKernel (/root/reference/PVCNN/modules/functional/src/knn/knn.cu:5-49,
// knn.cpp:6-25): dist [B,k,n] ascending, idx [B,k,n], slots (10000.0f, 0) when fewer than k references lie within
// d^2 < 10000, equal distances ordered by ascending reference index.  The reference scans all m references per query
// (O(n m k) with its list in global memory — seconds for a 50k-point scan); here each query looks only at the cells
// around it, and the result is still BIT-IDENTICAL to the brute-force scan because
//   * every candidate distance is the same expression  fma(dz,dz, fma(dy,dy, dx*dx)),  d* = query - reference;
//   * the k best are selected by the lexicographic key (distance, reference index), which is exactly "first k of a
//     stable sort by distance" — the order candidates are met in no longer matters;
//   * the search stops only when the k-th key is strictly below a conservative lower bound of every unvisited cell.
//
// Pipeline per call (all clouds of the batch in each launch):
//   grid_setup    one CTA per cloud: bounding box of the references -> cubic cells, G = round(cbrt(m/2)) <= 64 per axis
//                 (about 16 points per occupied cell on surface-like scans); clears the cell histograms.
//   grid_count    cell of every reference / query (queries clamped into the grid), integer histogram.
//   grid_scan     per cloud exclusive scan of both histograms.
//   grid_scatter  counting sort: references -> float4 (x, y, z, original index) in cell order; queries -> order list.
//   grid_query    one thread per query IN CELL ORDER (a warp's queries share their neighbourhood, so the candidate
//                 loads are broadcast/L1 hits and the warp does not diverge much); sorted top-k in registers;
//                 Chebyshev rings of cells, z-runs of a ring are contiguous in the sorted array.
#include "ri_common.cuh"

namespace {

constexpr float kUndefDist = 10000.0f;
constexpr int kMaxG = 64;
constexpr int kSetupThreads = 1024;

struct GridHdr {               // per cloud, 16 ints at the head of its workspace slab
    float ox, oy, oz, h, inv_h;
    int gx, gy, gz;
    int pad[8];
};

__host__ __device__ inline int grid_g(int m)
{
    int g = 1;
    while ((long long)(g + 1) * (g + 1) * (g + 1) * 2 <= (long long)(m > 0 ? m : 1) && g < kMaxG) ++g;   // ~cbrt(m/2)
    return g;
}

struct GridWs {                // int32 offsets inside one cloud's slab
    int hdr, rstart, rcur, qstart, qcur, qorder, sorted, stride, cells_max;
};
__host__ __device__ inline GridWs grid_ws_layout(int n, int m)
{
    GridWs w;
    const int g = grid_g(m);
    w.cells_max = g * g * g;
    int o = 0;
    w.hdr = o; o += 16;
    w.rstart = o; o += w.cells_max + 1;
    w.rcur = o; o += w.cells_max;
    w.qstart = o; o += w.cells_max + 1;
    w.qcur = o; o += w.cells_max;
    w.qorder = o; o += n;
    o = (o + 3) / 4 * 4;
    w.sorted = o; o += 4 * m;                                   // float4 per reference
    w.stride = (o + 3) / 4 * 4;
    return w;
}

__device__ __forceinline__ int cell_of(const GridHdr& H, float x, float y, float z, int& cx, int& cy, int& cz)
{
    cx = min(max((int)floorf((x - H.ox) * H.inv_h), 0), H.gx - 1);
    cy = min(max((int)floorf((y - H.oy) * H.inv_h), 0), H.gy - 1);
    cz = min(max((int)floorf((z - H.oz) * H.inv_h), 0), H.gz - 1);
    return (cx * H.gy + cy) * H.gz + cz;
}

__global__ void __launch_bounds__(kSetupThreads)
grid_setup_kernel(const float* __restrict__ refs, int m, int n, int* __restrict__ ws)
{
    __shared__ float smin[3][kSetupThreads / 32], smax[3][kSetupThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const GridWs L = grid_ws_layout(n, m);
    int* W = ws + (size_t)b * L.stride;
    const float* R = refs + (size_t)b * 3 * m;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = tid; j < m; j += kSetupThreads)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = R[j + (size_t)a * m];
            lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((tid & 31) == 0) { smin[a][tid >> 5] = lo[a]; smax[a][tid >> 5] = hi[a]; }
    }
    __syncthreads();
    if (tid == 0) {
        GridHdr H;
        float e[3];
        for (int a = 0; a < 3; ++a) {
            float l = smin[a][0], h = smax[a][0];
            for (int w = 1; w < kSetupThreads / 32; ++w) { l = fminf(l, smin[a][w]); h = fmaxf(h, smax[a][w]); }
            if (!(l <= h)) { l = 0.f; h = 0.f; }               // empty or NaN cloud
            (a == 0 ? H.ox : a == 1 ? H.oy : H.oz) = l;
            e[a] = h - l;
        }
        const int g = grid_g(m);
        const float emax = fmaxf(e[0], fmaxf(e[1], e[2]));
        H.h = emax > 0.f ? emax / (float)g : 1.0f;
        H.inv_h = 1.0f / H.h;
        H.gx = min(g, (int)(e[0] * H.inv_h) + 1);
        H.gy = min(g, (int)(e[1] * H.inv_h) + 1);
        H.gz = min(g, (int)(e[2] * H.inv_h) + 1);
        *reinterpret_cast<GridHdr*>(W + L.hdr) = H;
    }
    for (int c = tid; c < L.cells_max; c += kSetupThreads) { W[L.rcur + c] = 0; W[L.qcur + c] = 0; }
}

// pass 0: histogram into *cur; pass 1: scatter using *start + atomic cursor
template <int PASS>
__global__ void __launch_bounds__(256)
grid_bin_kernel(const float* __restrict__ queries, const float* __restrict__ refs, int n, int m, int* __restrict__ ws)
{
    const int b = blockIdx.y;
    const GridWs L = grid_ws_layout(n, m);
    int* W = ws + (size_t)b * L.stride;
    const GridHdr H = *reinterpret_cast<const GridHdr*>(W + L.hdr);
    const int t = blockIdx.x * 256 + threadIdx.x;
    int cx, cy, cz;
    if (t < m) {
        const float* R = refs + (size_t)b * 3 * m;
        const float x = R[t], y = R[t + m], z = R[t + 2 * (size_t)m];
        const int c = cell_of(H, x, y, z, cx, cy, cz);
        if (PASS == 0) atomicAdd(W + L.rcur + c, 1);
        else {
            const int pos = W[L.rstart + c] + atomicAdd(W + L.rcur + c, 1);
            reinterpret_cast<float4*>(W + L.sorted)[pos] = make_float4(x, y, z, __int_as_float(t));
        }
    }
    if (t < n) {
        const float* Q = queries + (size_t)b * 3 * n;
        const int c = cell_of(H, Q[t], Q[t + n], Q[t + 2 * (size_t)n], cx, cy, cz);
        if (PASS == 0) atomicAdd(W + L.qcur + c, 1);
        else W[L.qorder + W[L.qstart + c] + atomicAdd(W + L.qcur + c, 1)] = t;
    }
}

// exclusive scan of the two histograms of a cloud; the cursors are reset to 0 for the scatter pass
__global__ void __launch_bounds__(1024)
grid_scan_kernel(int n, int m, int* __restrict__ ws)
{
    __shared__ int swarp[32];
    __shared__ int scarry, stotal;
    const int b = blockIdx.x, which = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const GridWs L = grid_ws_layout(n, m);
    int* W = ws + (size_t)b * L.stride;
    int* cur = W + (which ? L.qcur : L.rcur);
    int* start = W + (which ? L.qstart : L.rstart);
    if (tid == 0) scarry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < L.cells_max; c0 += 1024) {
        const int c = c0 + tid;
        const int v = c < L.cells_max ? cur[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int w = swarp[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            swarp[lane] = wi - w;                               // exclusive offset of each warp
            if (lane == 31) stotal = wi;
        }
        __syncthreads();
        if (c < L.cells_max) { start[c] = scarry + swarp[wid] + incl - v; cur[c] = 0; }
        __syncthreads();
        if (tid == 0) scarry += stotal;
        __syncthreads();
    }
    if (tid == 0) start[L.cells_max] = scarry;
}

template <int KCAP>
struct KeyTopK {
    float d[KCAP];
    int j[KCAP];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int q = 0; q < KCAP; ++q) { d[q] = kUndefDist; j[q] = 0; }
    }
    static __device__ __forceinline__ bool before(float x, int xi, float y, int yi)
    {
        return x < y || (x == y && xi < yi);
    }
    // caller guarantees x < 10000 (so it sorts before every sentinel slot) and (x, idx) before slot KCAP-1
    __device__ __forceinline__ void insert(float x, int idx)
    {
#pragma unroll
        for (int q = KCAP - 1; q > 0; --q) {
            const bool up = before(x, idx, d[q - 1], j[q - 1]);
            const bool here = before(x, idx, d[q], j[q]);
            d[q] = up ? d[q - 1] : (here ? x : d[q]);
            j[q] = up ? j[q - 1] : (here ? idx : j[q]);
        }
        const bool first = before(x, idx, d[0], j[0]);
        d[0] = first ? x : d[0];
        j[0] = first ? idx : j[0];
    }
};

template <int KCAP>
__global__ void __launch_bounds__(128)
grid_query_kernel(const float* __restrict__ queries, int n, int m, int k, const int* __restrict__ ws,
                  float* __restrict__ dist, int* __restrict__ idx)
{
    const int b = blockIdx.y;
    const GridWs L = grid_ws_layout(n, m);
    const int* W = ws + (size_t)b * L.stride;
    const GridHdr H = *reinterpret_cast<const GridHdr*>(W + L.hdr);
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= n) return;
    const int qi = W[L.qorder + t];
    const float* Q = queries + (size_t)b * 3 * n;
    const float qx = Q[qi], qy = Q[qi + n], qz = Q[qi + 2 * (size_t)n];
    int cx, cy, cz;
    cell_of(H, qx, qy, qz, cx, cy, cz);
    const int* start = W + L.rstart;
    const float4* P = reinterpret_cast<const float4*>(W + L.sorted);

    KeyTopK<KCAP> top;
    top.init();
    const int rmax = max(max(cx, H.gx - 1 - cx), max(max(cy, H.gy - 1 - cy), max(cz, H.gz - 1 - cz)));
    for (int R = 0; R <= rmax; ++R) {
        const int x0 = max(cx - R, 0), x1 = min(cx + R, H.gx - 1);
        const int y0 = max(cy - R, 0), y1 = min(cy + R, H.gy - 1);
        for (int x = x0; x <= x1; ++x) {
            const bool xface = (x == cx - R) || (x == cx + R);
            for (int y = y0; y <= y1; ++y) {
                const bool face = xface || (y == cy - R) || (y == cy + R);
                // on a face of the ring the whole z-run belongs to it; inside, only the two end cells do
                const int nseg = (face || R == 0) ? 1 : 2;
                for (int sgi = 0; sgi < nseg; ++sgi) {
                    int z0, z1;
                    if (face || R == 0) { z0 = max(cz - R, 0); z1 = min(cz + R, H.gz - 1); }
                    else { z0 = z1 = (sgi == 0 ? cz - R : cz + R); if (z0 < 0 || z0 >= H.gz) continue; }
                    const int base = (x * H.gy + y) * H.gz;
                    const int p0 = __ldg(start + base + z0), p1 = __ldg(start + base + z1 + 1);
                    for (int p = p0; p < p1; ++p) {
                        const float4 r = __ldg(P + p);
                        const float dx = __fsub_rn(qx, r.x), dy = __fsub_rn(qy, r.y), dz = __fsub_rn(qz, r.z);
                        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                        const int j = __float_as_int(r.w);
                        if (d < kUndefDist && KeyTopK<KCAP>::before(d, j, top.d[KCAP - 1], top.j[KCAP - 1]))
                            top.insert(d, j);
                    }
                }
            }
        }
        // lower bound of the distance to anything outside the visited block (faces on the grid border: nothing beyond)
        float lb = INFINITY;
        if (cx - R > 0) lb = fminf(lb, qx - (H.ox + (float)(cx - R) * H.h));
        if (cx + R < H.gx - 1) lb = fminf(lb, (H.ox + (float)(cx + R + 1) * H.h) - qx);
        if (cy - R > 0) lb = fminf(lb, qy - (H.oy + (float)(cy - R) * H.h));
        if (cy + R < H.gy - 1) lb = fminf(lb, (H.oy + (float)(cy + R + 1) * H.h) - qy);
        if (cz - R > 0) lb = fminf(lb, qz - (H.oz + (float)(cz - R) * H.h));
        if (cz + R < H.gz - 1) lb = fminf(lb, (H.oz + (float)(cz + R + 1) * H.h) - qz);
        lb = fmaxf(lb - 1e-3f * H.h, 0.f);                      // cell assignment and the bound itself are rounded
        const float bound = lb * lb * 0.999f;
        if (top.d[KCAP - 1] < bound || bound >= kUndefDist) break;   // k-th key strictly inside, or nothing insertable left
    }
    float* od = dist + (size_t)b * k * n + qi;
    int* oi = idx + (size_t)b * k * n + qi;
#pragma unroll
    for (int s = 0; s < KCAP; ++s)
        if (s < k) { od[(size_t)s * n] = top.d[s]; oi[(size_t)s * n] = top.j[s]; }
}

}  // namespace

extern "C" size_t ri_knn_grid_workspace_bytes(int B, int n, int m)
{
    if (B <= 0 || n < 0 || m < 0) return 16;
    return (size_t)B * grid_ws_layout(n, m).stride * sizeof(int) + 16;
}

extern "C" int ri_knn_grid_f32(const float* xyz1, const float* xyz2, int B, int n, int m, int k,
                               float* dist1, int* idx1, void* workspace, size_t workspace_bytes, void* stream)
{
    if (B < 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (k > 32 || B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || n == 0) return RI_OK;
    if (workspace == nullptr || ((uintptr_t)workspace & 15) != 0 ||
        workspace_bytes < (size_t)B * grid_ws_layout(n, m).stride * sizeof(int)) return RI_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* ws = reinterpret_cast<int*>(workspace);
    grid_setup_kernel<<<B, kSetupThreads, 0, st>>>(xyz2, m, n, ws);
    RI_LAUNCH_CHECK();
    const int big = n > m ? n : m;
    dim3 gb((big + 255) / 256, B);
    grid_bin_kernel<0><<<gb, 256, 0, st>>>(xyz1, xyz2, n, m, ws);
    RI_LAUNCH_CHECK();
    grid_scan_kernel<<<dim3(B, 2), 1024, 0, st>>>(n, m, ws);
    RI_LAUNCH_CHECK();
    grid_bin_kernel<1><<<gb, 256, 0, st>>>(xyz1, xyz2, n, m, ws);
    RI_LAUNCH_CHECK();
    dim3 gq((n + 127) / 128, B);
    if (k <= 8) grid_query_kernel<8><<<gq, 128, 0, st>>>(xyz1, n, m, k, ws, dist1, idx1);
    else if (k <= 16) grid_query_kernel<16><<<gq, 128, 0, st>>>(xyz1, n, m, k, ws, dist1, idx1);
    else if (k <= 20) grid_query_kernel<20><<<gq, 128, 0, st>>>(xyz1, n, m, k, ws, dist1, idx1);
    else grid_query_kernel<32><<<gq, 128, 0, st>>>(xyz1, n, m, k, ws, dist1, idx1);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
