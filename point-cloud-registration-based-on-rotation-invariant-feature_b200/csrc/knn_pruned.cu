// knn_pruned.cu — exact k-nearest-neighbour search for small clouds (n, m <= 2048, c == 3, k <= 32) that scans
// only the part of the reference set a query can still improve on.  sm_100a.
//
// Same contract as knn.cu (KnnKernel, /root/reference/PVCNN/modules/functional/src/knn/knn.cu:5-49): squared L2
// d = fma(dz,dz,fma(dy,dy,dx*dx)) with d_p = query_p - ref_p, slots start as (10000.0f, 0), a candidate enters only
// below the current k-th entry, result = the first k of a stable sort by distance (the lower reference index wins
// ties).  The brute-force kernel gets that order for free by visiting the references in index order; here the visiting
// order is spatial, so the list is kept sorted by the KEY (distance, reference index) — the same total order, whatever
// the order of arrival.
//
//   knn_prep_kernel   one CTA per (cloud, point set): 12-bit Morton cell of every point (4 bits per axis over the set's
//                     bounding box), counting sort by cell in shared memory, points written out in that order as
//                     float4 (x, y, z, original index), cut into 32 blocks of BS = 32 or 64 consecutive points, each
//                     with its axis-aligned bounding box.  (Order inside a cell follows the shared-memory atomics: it
//                     only decides which block a point lands in, never a result.)
//   knn3_pruned_kernel  one thread per query, queries taken in their own Morton order so the 32 queries of a warp are
//                     neighbours in space.  The warp orders the reference blocks by the distance between its queries'
//                     bounding box and the block's (lane b keeps the bound of block b; the next block is a
//                     redux.sync.min), stops at the first block whose bound exceeds every lane's current k-th distance,
//                     skips blocks no lane can improve on (per-lane point-to-box bound), and runs the brute-force
//                     kernel's inner loop on the others: 32 candidates filtered against the k-th distance into a bit
//                     mask, survivors inserted into the register-resident sorted list.
// Both bounds are computed with the same rounding sequence as the distance itself (fsub, fmul, fma, fma — each monotone
// in |d_p|), so a bound never exceeds the distance the kernel would have computed for any point inside the box: the
// pruning cannot change a result bit.  NaN coordinates make every comparison false: such points never enter, such
// queries keep their default slots — as in the reference.
#include "ri_common.cuh"

namespace {

constexpr float kUndefDist = 10000.0f;   // knn/knn.cuh:3 (UNDEFINE_VALUE)
constexpr int kMaxPts = 2048;            // largest point set of this path
constexpr int kBlocks = 32;              // reference blocks per cloud: one per lane
constexpr int kBoxF4 = 2 * kBlocks;      // float4 (lo), float4 (hi) per block
constexpr int kCells = 4096;             // 4 Morton bits per axis
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int block_size_for(int m) { return m > 1024 ? 64 : 32; }   // points per block (m <= 2048)
// float4 elements one sorted set occupies in the workspace: 32 blocks of points, then the boxes
__host__ __device__ inline size_t set_stride4(int m) { return (size_t)kBlocks * block_size_for(m) + kBoxF4; }

__device__ __forceinline__ unsigned spread4(unsigned v)
{
    v = (v | (v << 4)) & 0x0C3u;
    v = (v | (v << 2)) & 0x249u;
    return v;
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------ prep
// grid (B, sets), 1024 threads.  set 0 = the points of `pa` ([B,3,na]), set 1 = those of `pb` ([B,3,nb]).
__global__ void __launch_bounds__(1024)
knn_prep_kernel(const float* __restrict__ pa, int na, const float* __restrict__ pb, int nb,
                float4* __restrict__ wsa, float4* __restrict__ wsb)
{
    extern __shared__ float4 sorted[];               // kBlocks * BS
    __shared__ int hist[kCells];
    __shared__ float red[6][32];
    __shared__ int wtot[32];

    const int b = blockIdx.x;
    const bool second = blockIdx.y != 0;
    const int cnt = second ? nb : na;
    const float* P = (second ? pb : pa) + (size_t)b * 3 * cnt;
    float4* W = (second ? wsb : wsa) + (size_t)b * set_stride4(cnt);
    const int BS = block_size_for(cnt);
    const int nblk = (cnt + BS - 1) / BS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    float x[2], y[2], z[2];
    bool v[2];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * 1024;
        v[u] = i < cnt;
        x[u] = y[u] = z[u] = 0.f;
        if (v[u]) {
            x[u] = P[i]; y[u] = P[i + cnt]; z[u] = P[i + 2 * (size_t)cnt];
            // finite values only: an infinite coordinate would blow the cell size up for everybody else
            if (fabsf(x[u]) <= 3.0e38f) { lo[0] = fminf(lo[0], x[u]); hi[0] = fmaxf(hi[0], x[u]); }
            if (fabsf(y[u]) <= 3.0e38f) { lo[1] = fminf(lo[1], y[u]); hi[1] = fmaxf(hi[1], y[u]); }
            if (fabsf(z[u]) <= 3.0e38f) { lo[2] = fminf(lo[2], z[u]); hi[2] = fmaxf(hi[2], z[u]); }
        }
    }
    for (int t = tid; t < kCells; t += 1024) hist[t] = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min(lo[a]); hi[a] = warp_max(hi[a]);
        if (lane == 0) { red[a][wid] = lo[a]; red[3 + a][wid] = hi[a]; }
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float l = warp_min(red[a][lane]), h = warp_max(red[3 + a][lane]);
            if (lane == 0) { red[a][0] = l; red[3 + a][0] = h; }
        }
    }
    __syncthreads();
    float inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = red[a][0];
        const float ext = red[3 + a][0] - lo[a];
        inv[a] = (ext > 0.f && ext <= 3.0e38f) ? 16.0f / ext : 0.f;
    }
    unsigned code[2];
    int rank[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (v[u]) {
            // fmaxf(NaN, 0) == 0: a NaN / infinite coordinate lands in cell 0 or 15 of its axis
            const unsigned cx = (unsigned)fminf(fmaxf((x[u] - lo[0]) * inv[0], 0.f), 15.f);
            const unsigned cy = (unsigned)fminf(fmaxf((y[u] - lo[1]) * inv[1], 0.f), 15.f);
            const unsigned cz = (unsigned)fminf(fmaxf((z[u] - lo[2]) * inv[2], 0.f), 15.f);
            code[u] = spread4(cx) | (spread4(cy) << 1) | (spread4(cz) << 2);
            rank[u] = atomicAdd(&hist[code[u]], 1);
        }
    }
    __syncthreads();
    {   // exclusive scan of the 4096 cell counts: 4 consecutive bins per thread
        const int c0 = hist[4 * tid], c1 = hist[4 * tid + 1], c2 = hist[4 * tid + 2], c3 = hist[4 * tid + 3];
        const int s = c0 + c1 + c2 + c3;
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = wtot[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, wi, o);
                if (lane >= o) wi += t;
            }
            wtot[lane] = wi - w;
        }
        __syncthreads();
        const int base = wtot[wid] + inc - s;
        hist[4 * tid] = base; hist[4 * tid + 1] = base + c0; hist[4 * tid + 2] = base + c0 + c1;
        hist[4 * tid + 3] = base + c0 + c1 + c2;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (v[u]) sorted[hist[code[u]] + rank[u]] = make_float4(x[u], y[u], z[u], __int_as_float(tid + u * 1024));
    for (int t = cnt + tid; t < nblk * BS; t += 1024)      // tail of the last block: points no query can reach
        sorted[t] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
    __syncthreads();
    for (int t = tid; t < nblk * BS; t += 1024) W[t] = sorted[t];
    // bounding box of block `wid`
    float4 blo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), bhi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    if (wid < nblk) {
        for (int e = lane; e < BS; e += 32) {
            const int t = wid * BS + e;
            if (t < cnt) {
                const float4 p = sorted[t];
                blo.x = fminf(blo.x, p.x); blo.y = fminf(blo.y, p.y); blo.z = fminf(blo.z, p.z);
                bhi.x = fmaxf(bhi.x, p.x); bhi.y = fmaxf(bhi.y, p.y); bhi.z = fmaxf(bhi.z, p.z);
            }
        }
        blo.x = warp_min(blo.x); blo.y = warp_min(blo.y); blo.z = warp_min(blo.z);
        bhi.x = warp_max(bhi.x); bhi.y = warp_max(bhi.y); bhi.z = warp_max(bhi.z);
    }
    if (lane == 0) {
        float4* box = W + (size_t)kBlocks * BS;
        box[2 * wid] = blo; box[2 * wid + 1] = bhi;
    }
}

// ------------------------------------------------------------------------------------------------ search
template <int KCAP>
struct KeyedTopK {
    // slot = (distance bits << 32) | reference index: distances are >= +0, so the unsigned order of the bits is the
    // order of the values, and the 64-bit order is (distance, index)
    unsigned long long s[KCAP];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int q = 0; q < KCAP; ++q) s[q] = (unsigned long long)__float_as_uint(kUndefDist) << 32;
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float((unsigned)(s[KCAP - 1] >> 32)); }
    __device__ __forceinline__ unsigned long long worst_key() const { return s[KCAP - 1]; }
    __device__ __forceinline__ void insert(unsigned long long x)   // caller guarantees x < s[KCAP-1]
    {
#pragma unroll
        for (int q = KCAP - 1; q > 0; --q) {
            const bool up = x < s[q - 1];
            const bool here = x < s[q];
            s[q] = up ? s[q - 1] : (here ? x : s[q]);
        }
        s[0] = x < s[0] ? x : s[0];
    }
};

__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, const float4 r)
{
    const float dx = __fsub_rn(qx, r.x), dy = __fsub_rn(qy, r.y), dz = __fsub_rn(qz, r.z);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
// gap between an interval [lo, hi] and a point / interval, rounded like the kernel's own subtraction
__device__ __forceinline__ float gap1(float lo, float hi, float qlo, float qhi)
{
    return fmaxf(fmaxf(__fsub_rn(lo, qhi), __fsub_rn(qlo, hi)), 0.0f);
}
__device__ __forceinline__ float bound3(float gx, float gy, float gz)
{
    return __fmaf_rn(gz, gz, __fmaf_rn(gy, gy, __fmul_rn(gx, gx)));
}

// refs / queries: sorted sets written by knn_prep_kernel (cloud b at + b * stride4).  self != 0: the queries are the
// references (same set), read from shared memory.
template <int KCAP, int QPC>
__global__ void __launch_bounds__(QPC)
knn3_pruned_kernel(const float4* __restrict__ wref, const float4* __restrict__ wqry, int self, int n, int m, int k,
                   float* __restrict__ dist, int* __restrict__ idx)
{
    extern __shared__ float4 sm[];                   // nblk * BS reference points, then the 64 box entries
    const int b = blockIdx.y;
    const int BS = block_size_for(m);
    const int nblk = (m + BS - 1) / BS;
    const int mp = nblk * BS;
    const float4* R = wref + (size_t)b * set_stride4(m);
    float4* sbox = sm + mp;
    for (int t = threadIdx.x; t < mp; t += QPC) sm[t] = R[t];
    for (int t = threadIdx.x; t < kBoxF4; t += QPC) sbox[t] = R[(size_t)kBlocks * BS + t];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * QPC + threadIdx.x;
    const bool live = p < n;
    const int pq = live ? p : n - 1;
    const float4 q4 = self ? sm[pq] : wqry[(size_t)b * set_stride4(n) + pq];
    const float qx = q4.x, qy = q4.y, qz = q4.z;

    // warp-level order of the blocks: bound between the warp's query box and each block's box
    unsigned key = 0xffffffffu;
    {
        const float wlx = warp_min(qx), wly = warp_min(qy), wlz = warp_min(qz);
        const float whx = warp_max(qx), why = warp_max(qy), whz = warp_max(qz);
        if (lane < nblk) {
            const float4 blo = sbox[2 * lane], bhi = sbox[2 * lane + 1];
            const float lb = bound3(gap1(blo.x, bhi.x, wlx, whx), gap1(blo.y, bhi.y, wly, why), gap1(blo.z, bhi.z, wlz, whz));
            key = (__float_as_uint(lb) & ~31u) | (unsigned)lane;     // lb >= 0; low mantissa bits dropped: still a lower bound
        }
    }

    KeyedTopK<KCAP> top;
    top.init();
    while (true) {
        const unsigned kmin = __reduce_min_sync(kFull, key);
        if (kmin == 0xffffffffu) break;
        const int blk = kmin & 31;
        const float wbound = __uint_as_float(kmin & ~31u);
        const float wworst = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(top.worst())));
        if (wbound > wworst) break;                  // every block still waiting is at least this far from every lane
        if (lane == blk) key = 0xffffffffu;
        {
            const float4 blo = sbox[2 * blk], bhi = sbox[2 * blk + 1];
            const float lb = bound3(gap1(blo.x, bhi.x, qx, qx), gap1(blo.y, bhi.y, qy, qy), gap1(blo.z, bhi.z, qz, qz));
            if (!__any_sync(kFull, lb <= top.worst())) continue;
        }
        for (int h = 0; h < BS; h += 32) {
            const float4* cand = sm + blk * BS + h;
            const float thr = top.worst();
            unsigned mask = 0u;
#pragma unroll
            for (int u = 0; u < 32; ++u)
                mask |= (sqdist3(qx, qy, qz, cand[u]) <= thr) ? (1u << u) : 0u;
            while (mask) {
                const int u = __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 r = cand[u];
                const unsigned long long x = ((unsigned long long)__float_as_uint(sqdist3(qx, qy, qz, r)) << 32) |
                                             (unsigned)__float_as_int(r.w);
                if (x < top.worst_key()) top.insert(x);
            }
        }
    }
    if (live) {
        const int qi = __float_as_int(q4.w);
        float* od = dist + (size_t)b * k * n + qi;
        int* oi = idx + (size_t)b * k * n + qi;
#pragma unroll
        for (int s = 0; s < KCAP; ++s)
            if (s < k) {
                od[(size_t)s * n] = __uint_as_float((unsigned)(top.s[s] >> 32));
                oi[(size_t)s * n] = (int)(unsigned)(top.s[s] & 0xffffffffu);
            }
    }
}

int qpc_setting()
{
    static int v = 0;
    if (v == 0) {
        v = 64;
        if (const char* ev = getenv("RI_KNN_QPC")) { const int t = atoi(ev); if (t == 32 || t == 64 || t == 128) v = t; }
    }
    return v;
}

template <int KCAP, int QPC>
void launch_search_q(const float4* wref, const float4* wqry, int self, int B, int n, int m, int k, float* dist, int* idx,
                     cudaStream_t st)
{
    const int BS = block_size_for(m);
    const size_t smem = ((size_t)((m + BS - 1) / BS) * BS + kBoxF4) * sizeof(float4);
    cudaFuncSetAttribute(knn3_pruned_kernel<KCAP, QPC>, cudaFuncAttributePreferredSharedMemoryCarveout, ri_step_carveout_percent());
    dim3 grid((n + QPC - 1) / QPC, B);
    knn3_pruned_kernel<KCAP, QPC><<<grid, QPC, smem, st>>>(wref, wqry, self, n, m, k, dist, idx);
}

template <int KCAP>
void launch_search_k(const float4* wref, const float4* wqry, int self, int B, int n, int m, int k, float* dist, int* idx,
                     cudaStream_t st)
{
    switch (qpc_setting()) {
    case 32: launch_search_q<KCAP, 32>(wref, wqry, self, B, n, m, k, dist, idx, st); break;
    case 128: launch_search_q<KCAP, 128>(wref, wqry, self, B, n, m, k, dist, idx, st); break;
    default: launch_search_q<KCAP, 64>(wref, wqry, self, B, n, m, k, dist, idx, st); break;
    }
}

int launch_search(const float4* wref, const float4* wqry, int self, int B, int n, int m, int k, float* dist, int* idx,
                  cudaStream_t st)
{
    if (k <= 8) launch_search_k<8>(wref, wqry, self, B, n, m, k, dist, idx, st);
    else if (k <= 16) launch_search_k<16>(wref, wqry, self, B, n, m, k, dist, idx, st);
    else if (k <= 20) launch_search_k<20>(wref, wqry, self, B, n, m, k, dist, idx, st);
    else launch_search_k<32>(wref, wqry, self, B, n, m, k, dist, idx, st);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

int launch_prep(const float* pa, int na, const float* pb, int nb, int B, float4* wsa, float4* wsb, cudaStream_t st)
{
    const int big = na > nb ? na : nb;
    const size_t smem = (size_t)kBlocks * block_size_for(big) * sizeof(float4);
    cudaFuncSetAttribute(knn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBlocks * 64 * sizeof(float4)));
    cudaFuncSetAttribute(knn_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ri_step_carveout_percent());
    knn_prep_kernel<<<dim3(B, pb != nullptr ? 2 : 1), 1024, smem, st>>>(pa, na, pb, nb, wsa, wsb);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

inline bool pruned_ok(int c, int n, int m, int k)
{
    return c == 3 && k <= 32 && n >= 1 && m >= 1 && n <= kMaxPts && m <= kMaxPts;
}

}  // namespace

extern "C" int ri_knn_f32(const float*, const float*, int, int, int, int, int, float*, int*, void*);

// ---- C ABI (include/ri_b200.h) ----------------------------------------------------------------------
extern "C" size_t ri_knn_workspace_bytes(int B, int n, int m)
{
    if (B <= 0 || n <= 0 || m <= 0 || n > kMaxPts || m > kMaxPts) return 16;
    return (size_t)B * (set_stride4(n) + set_stride4(m)) * sizeof(float4);
}

extern "C" int ri_knn_ws_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                             float* dist1, int* idx1, void* workspace, size_t workspace_bytes, void* stream)
{
    if (B < 0 || c <= 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (!pruned_ok(c, n, m, k)) return ri_knn_f32(xyz1, xyz2, B, c, n, m, k, dist1, idx1, stream);
    if (B == 0) return RI_OK;
    if (workspace == nullptr || workspace_bytes < ri_knn_workspace_bytes(B, n, m) || ((uintptr_t)workspace & 15)) return RI_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float4* wref = (float4*)workspace;
    const bool self = xyz1 == xyz2 && n == m;
    float4* wqry = self ? wref : wref + (size_t)B * set_stride4(m);
    int rc = launch_prep(xyz2, m, self ? nullptr : xyz1, n, B, wref, wqry, st);
    if (rc != RI_OK) return rc;
    return launch_search(wref, wqry, self ? 1 : 0, B, n, m, k, dist1, idx1, st);
}

extern "C" int ri_knn_bilateral_ws_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                                       float* dist1, float* dist2, int* idx1, int* idx2,
                                       void* workspace, size_t workspace_bytes, void* stream)
{
    if (B < 0 || c <= 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (!pruned_ok(c, n, m, k)) {
        int rc = ri_knn_f32(xyz1, xyz2, B, c, n, m, k, dist1, idx1, stream);
        if (rc != RI_OK) return rc;
        return ri_knn_f32(xyz2, xyz1, B, c, m, n, k, dist2, idx2, stream);
    }
    if (B == 0) return RI_OK;
    if (workspace == nullptr || workspace_bytes < ri_knn_workspace_bytes(B, n, m) || ((uintptr_t)workspace & 15)) return RI_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float4* w2 = (float4*)workspace;                          // sorted xyz2
    float4* w1 = w2 + (size_t)B * set_stride4(m);             // sorted xyz1
    const bool self = xyz1 == xyz2 && n == m;
    int rc = launch_prep(xyz2, m, self ? nullptr : xyz1, n, B, w2, w1, st);
    if (rc != RI_OK) return rc;
    if (self) {
        rc = launch_search(w2, w2, 1, B, n, m, k, dist1, idx1, st);
        if (rc != RI_OK) return rc;
        return launch_search(w2, w2, 1, B, m, n, k, dist2, idx2, st);
    }
    rc = launch_search(w2, w1, 0, B, n, m, k, dist1, idx1, st);
    if (rc != RI_OK) return rc;
    return launch_search(w1, w2, 0, B, m, n, k, dist2, idx2, st);
}
