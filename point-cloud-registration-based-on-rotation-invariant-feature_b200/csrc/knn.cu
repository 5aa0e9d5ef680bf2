// knn.cu — brute-force k-nearest-neighbour search for small clouds (N ~ 1024), sm_100a.
//
// Replaces  KnnKernel  (/root/reference/PVCNN/modules/functional/src/knn/knn.cu:5-49) and its host
// wrapper knn_forward_cuda (knn/knn.cpp:6-25).  Semantics kept bit-exactly:
//   * d = fma(d_p, d_p, acc) over channels p = 0..c-1 with d_p = query_p - ref_p, acc starts at 0;
//   * slots start as (10000.0f, 0); a candidate enters iff d < slot[k-1] (strict) and then bubbles up
//     past strictly larger entries only  =>  result = first k of a stable sort by distance, the lower
//     reference index wins ties, candidates with d >= 10000 (or NaN) never enter.
//
// B200 design (not the reference's one-CTA-per-cloud global-memory insertion sort):
//   * grid = (query tiles, clouds): a 1024-point cloud is spread over 8-16 CTAs so 32 clouds fill 148 SMs;
//   * the reference set is staged once per CTA into shared memory as float4 (x,y,z,-) — every lane of a
//     warp reads the same reference point, i.e. one broadcast LDS.128 per candidate;
//   * one thread owns one query and keeps its sorted top-k (distance, index) list in REGISTERS;
//   * candidates are filtered 32 at a time against the current k-th distance into a bit mask (no
//     divergence), and only the survivors are re-evaluated and inserted with a fully unrolled,
//     predicated shift — ~k*(1+ln(m/k)) insertions per query instead of m.
#include "ri_common.cuh"
#include "ppf_math.cuh"

int ri_launch_knn_warp(const float* queries, const float* refs, int B, int n, int m, int k, float* dist, int* idx,
                       cudaStream_t st);

namespace {

constexpr float kUndefDist = 10000.0f;   // knn/knn.cuh:3 (UNDEFINE_VALUE)
constexpr int kQueriesPerCta = 128;
constexpr int kRefTile = 2048;           // reference points staged per pass (32 KB of float4)

template <int KCAP>
struct TopK {
    float d[KCAP];
    int j[KCAP];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int q = 0; q < KCAP; ++q) { d[q] = kUndefDist; j[q] = 0; }
    }
    __device__ __forceinline__ float worst() const { return d[KCAP - 1]; }
    // Stable insert of (x, idx); caller guarantees x < d[KCAP-1].
    __device__ __forceinline__ void insert(float x, int idx)
    {
#pragma unroll
        for (int q = KCAP - 1; q > 0; --q) {
            const bool up = x < d[q - 1];      // x sorts strictly before slot q-1: slot q inherits q-1
            const bool here = x < d[q];        // x sorts before the old slot q
            d[q] = up ? d[q - 1] : (here ? x : d[q]);
            j[q] = up ? j[q - 1] : (here ? idx : j[q]);
        }
        const bool first = x < d[0];
        d[0] = first ? x : d[0];
        j[0] = first ? idx : j[0];
    }
};

__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, const float4 r)
{
    const float dx = __fsub_rn(qx, r.x), dy = __fsub_rn(qy, r.y), dz = __fsub_rn(qz, r.z);
    // fma(dx,dx,0) == dx*dx rounded once: identical to the reference's first chain link.
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// c == 3 fast path.  queries [B,3,n], refs [B,3,m] -> dist/idx [B,k,n]
template <int KCAP>
__global__ void __launch_bounds__(kQueriesPerCta)
knn3_kernel(const float* __restrict__ queries, const float* __restrict__ refs, int n, int m, int k,
            float* __restrict__ dist, int* __restrict__ idx)
{
    extern __shared__ float4 sref[];                 // min(m, kRefTile) reference points
    const int b = blockIdx.y;
    const int i = blockIdx.x * kQueriesPerCta + threadIdx.x;
    const bool live = i < n;
    const float* q = queries + (size_t)b * 3 * n;
    const float* rf = refs + (size_t)b * 3 * m;
    const int iq = live ? i : n - 1;
    const float qx = q[iq], qy = q[iq + n], qz = q[iq + 2 * (size_t)n];

    TopK<KCAP> top;
    top.init();

    for (int base = 0; base < m; base += kRefTile) {
        const int cnt = min(kRefTile, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += kQueriesPerCta) {
            const int g = base + t;
            sref[t] = make_float4(rf[g], rf[g + m], rf[g + 2 * (size_t)m], 0.0f);
        }
        __syncthreads();
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            const int lim = min(32, cnt - c0);
            const float thr = top.worst();
            unsigned mask = 0u;
            if (lim == 32) {
#pragma unroll
                for (int u = 0; u < 32; ++u)
                    mask |= (sqdist3(qx, qy, qz, sref[c0 + u]) < thr) ? (1u << u) : 0u;
            } else {
                for (int u = 0; u < lim; ++u)
                    mask |= (sqdist3(qx, qy, qz, sref[c0 + u]) < thr) ? (1u << u) : 0u;
            }
            while (mask) {                      // ascending reference index: keeps the stable order
                const int u = __ffs(mask) - 1;
                mask &= mask - 1;
                const float d = sqdist3(qx, qy, qz, sref[c0 + u]);
                if (d < top.worst()) top.insert(d, base + c0 + u);
            }
        }
    }
    if (live) {
        float* od = dist + (size_t)b * k * n + i;
        int* oi = idx + (size_t)b * k * n + i;
#pragma unroll
        for (int s = 0; s < KCAP; ++s)
            if (s < k) { od[(size_t)s * n] = top.d[s]; oi[(size_t)s * n] = top.j[s]; }
    }
}

// Fused self-query k-NN + point-pair features.  Same search as knn3_kernel (queries == references, whole cloud staged
// in shared memory), and since the thread that owns a query ends up holding its k neighbours in registers, it evaluates
// the PPF columns (centre = the query, point = each neighbour; ppf_math.cuh) right there: the neighbour's coordinates are
// already in shared memory, its normal is staged next to them, and the [B,k,N] index tensor is never re-read.  Equals
// ri_knn_f32(xyz, xyz) followed by ri_ppf_gather_f32 bit for bit.
template <int KCAP>
__global__ void __launch_bounds__(kQueriesPerCta)
knn3_ppf_kernel(const float* __restrict__ xyz, const float* __restrict__ normals, long long cloud_stride, int n, int k,
                float* __restrict__ dist, int* __restrict__ idx, float* __restrict__ ppf)
{
    extern __shared__ float4 sref[];                 // [n] (x, y, z, nx) then [n] (ny, nz, -, -)
    float4* snrm = sref + n;
    const int b = blockIdx.y;
    const int i = blockIdx.x * kQueriesPerCta + threadIdx.x;
    const bool live = i < n;
    const float* X = xyz + (size_t)b * cloud_stride;
    const float* Nn = normals + (size_t)b * cloud_stride;
    const size_t n2 = 2 * (size_t)n;
    for (int t = threadIdx.x; t < n; t += kQueriesPerCta) {
        sref[t] = make_float4(X[t], X[t + n], X[t + n2], Nn[t]);
        snrm[t] = make_float4(Nn[t + n], Nn[t + n2], 0.f, 0.f);
    }
    __syncthreads();
    const int iq = live ? i : n - 1;
    const float4 me = sref[iq];
    const float qx = me.x, qy = me.y, qz = me.z;

    TopK<KCAP> top;
    top.init();
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int lim = min(32, n - c0);
        const float thr = top.worst();
        unsigned mask = 0u;
        if (lim == 32) {
#pragma unroll
            for (int u = 0; u < 32; ++u)
                mask |= (sqdist3(qx, qy, qz, sref[c0 + u]) < thr) ? (1u << u) : 0u;
        } else {
            for (int u = 0; u < lim; ++u)
                mask |= (sqdist3(qx, qy, qz, sref[c0 + u]) < thr) ? (1u << u) : 0u;
        }
        while (mask) {                          // ascending reference index: keeps the stable order
            const int u = __ffs(mask) - 1;
            mask &= mask - 1;
            const float d = sqdist3(qx, qy, qz, sref[c0 + u]);
            if (d < top.worst()) top.insert(d, c0 + u);
        }
    }
    if (!live) return;
    const size_t kn = (size_t)k * n;
    if (dist != nullptr) {
        float* od = dist + (size_t)b * kn + i;
        int* oi = idx + (size_t)b * kn + i;
#pragma unroll
        for (int s = 0; s < KCAP; ++s)
            if (s < k) { od[(size_t)s * n] = top.d[s]; oi[(size_t)s * n] = top.j[s]; }
    }
    const float4 mn = snrm[i];
    float* O = ppf + (size_t)b * 4 * kn + i;
#pragma unroll
    for (int s = 0; s < KCAP; ++s) {
        if (s < k) {
            const int j = top.j[s];              // an unfilled slot holds index 0, exactly what the two-kernel path gathers
            const float4 p = sref[j], pq = snrm[j];
            const Ppf4 r = ppf_column(me.x, me.y, me.z, me.w, mn.x, mn.y, p.x, p.y, p.z, p.w, pq.x, pq.y);
            float* o = O + (size_t)s * n;
            o[0] = r.a1; o[kn] = r.a2; o[2 * kn] = r.a3; o[3 * kn] = r.dn;
        }
    }
}

// Generic path: any channel count c, any k.  Thread per query, list kept in the output arrays
// (coalesced along the query index).  Correct for everything; used only off the hot configuration.
__global__ void __launch_bounds__(kQueriesPerCta)
knn_generic_kernel(const float* __restrict__ queries, const float* __restrict__ refs, int c, int n, int m, int k,
                   float* __restrict__ dist, int* __restrict__ idx)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kQueriesPerCta + threadIdx.x;
    if (i >= n) return;
    const float* q = queries + (size_t)b * c * n;
    const float* rf = refs + (size_t)b * c * m;
    float* od = dist + (size_t)b * k * n + i;
    int* oi = idx + (size_t)b * k * n + i;
    for (int s = 0; s < k; ++s) { od[(size_t)s * n] = kUndefDist; oi[(size_t)s * n] = 0; }
    float worst = kUndefDist;
    for (int j = 0; j < m; ++j) {
        float d = 0.0f;
        for (int p = 0; p < c; ++p) {
            const float df = __fsub_rn(q[i + (size_t)p * n], rf[j + (size_t)p * m]);
            d = __fmaf_rn(df, df, d);
        }
        if (d < worst) {
            int s = k - 1;
            while (s > 0 && d < od[(size_t)(s - 1) * n]) {
                od[(size_t)s * n] = od[(size_t)(s - 1) * n];
                oi[(size_t)s * n] = oi[(size_t)(s - 1) * n];
                --s;
            }
            od[(size_t)s * n] = d;
            oi[(size_t)s * n] = j;
            worst = od[(size_t)(k - 1) * n];
        }
    }
}

int launch_knn(const float* queries, const float* refs, int B, int c, int n, int m, int k,
               float* dist, int* idx, cudaStream_t st, bool thread_per_query)
{
    if (B == 0 || n == 0) return RI_OK;
    if (!thread_per_query && c == 3 && k <= 32 && m >= 1 && m <= 1024)
        return ri_launch_knn_warp(queries, refs, B, n, m, k, dist, idx, st);      // knn_warp.cu: a warp per query
    dim3 grid((n + kQueriesPerCta - 1) / kQueriesPerCta, B);
    const size_t smem = (size_t)(m < kRefTile ? (m > 0 ? m : 1) : kRefTile) * sizeof(float4);
    if (c == 3 && k <= 32) {
        auto kern = k <= 8 ? knn3_kernel<8> : k <= 16 ? knn3_kernel<16> : k <= 20 ? knn3_kernel<20> : knn3_kernel<32>;
        RI_KERNEL_SETUP(kern, false, ri_step_carveout_percent());
        kern<<<grid, kQueriesPerCta, smem, st>>>(queries, refs, n, m, k, dist, idx);
    } else {
        knn_generic_kernel<<<grid, kQueriesPerCta, 0, st>>>(queries, refs, c, n, m, k, dist, idx);
    }
    RI_LAUNCH_CHECK();
    return RI_OK;
}

}  // namespace

// ---- C ABI (include/ri_b200.h) ----------------------------------------------------------------------
extern "C" int ri_knn_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                          float* dist1, int* idx1, void* stream)
{
    if (B < 0 || c <= 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    return launch_knn(xyz1, xyz2, B, c, n, m, k, dist1, idx1, (cudaStream_t)stream, false);
}

// The thread-per-query form for every shape (what ri_knn_f32 runs for m > 1024, c != 3 or k > 32): kept addressable so the
// tests can hold the two forms against each other.
extern "C" int ri_knn_thread_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                                 float* dist1, int* idx1, void* stream)
{
    if (B < 0 || c <= 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    return launch_knn(xyz1, xyz2, B, c, n, m, k, dist1, idx1, (cudaStream_t)stream, true);
}

extern "C" int ri_knn_bilateral_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                                    float* dist1, float* dist2, int* idx1, int* idx2, void* stream)
{
    int rc = ri_knn_f32(xyz1, xyz2, B, c, n, m, k, dist1, idx1, stream);
    if (rc != RI_OK) return rc;
    return ri_knn_f32(xyz2, xyz1, B, c, m, n, k, dist2, idx2, stream);
}

// Fused self-query k-NN + PPF (see knn3_ppf_kernel).  xyz / normals address [3,N] planes per cloud, cloud b at
// base + b * cloud_stride floats (3N for two contiguous [B,3,N] arrays; 6N with normals = xyz + 3N for the interleaved
// [B,6,N] input batch).  dist / idx may both be null when only the features are wanted.
extern "C" int ri_knn_ppf_f32(const float* xyz, const float* normals, long long cloud_stride, int B, int N, int k,
                              float* dist, int* idx, float* ppf, void* stream)
{
    if (B < 0 || N < 0 || k <= 0 || cloud_stride < 0 || ppf == nullptr || ((dist == nullptr) != (idx == nullptr)))
        return RI_ERR_BAD_ARG;
    if (k > 32 || N > kRefTile || B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || N == 0) return RI_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)N * 2 * sizeof(float4);
    dim3 grid((N + kQueriesPerCta - 1) / kQueriesPerCta, B);
    auto kern = k <= 8 ? knn3_ppf_kernel<8> : k <= 16 ? knn3_ppf_kernel<16> : k <= 20 ? knn3_ppf_kernel<20> : knn3_ppf_kernel<32>;
    RI_KERNEL_SETUP(kern, true, ri_step_carveout_percent());        // up to 64 KB of dynamic shared memory: per-device opt-in
    kern<<<grid, kQueriesPerCta, smem, st>>>(xyz, normals, cloud_stride, N, k, dist, idx, ppf);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
