// ballquery.cu — radius ("ball") query and neighbour grouping, sm_100a.   SURVEY.md §8(f) row f1.
//
// Replaces  ball_query_kernel   (/root/reference/PVCNN/modules/functional/src/ball_query/ball_query.cu:19-50,
//                                host ball_query.cpp:6-30)
//           grouping_kernel / grouping_grad_kernel  (.../src/grouping/grouping.cu:18-44, 58-84; host grouping.cpp)
// — the neighbourhood the shipped classification / registration models actually use for their local features
// (BallQuery(r = 0.3, u = 128), PVCNN/models/pvcnn_classify.py:61-67, 252-271).
//
// Ball query semantics kept bit-exactly: for centre j the points k = 0, 1, ... are scanned in index order;
// d2 = fma(dz,dz, fma(dy,dy, dx*dx)) with d* = centre - point (the contraction nvcc chose, read from the reference's
// sm_100a SASS); k is a neighbour iff d2 < r2 && (double)d2 > 1e-5 (the centre itself and near-duplicates are
// excluded); the FIRST neighbour fills all u slots, each further one overwrites slot cnt; the scan stops after u
// neighbours; a centre without neighbours keeps u zeros (the reference's torch::zeros).
//
// B200 design.  The reference runs one CTA per cloud with one THREAD per centre walking the whole cloud (32 CTAs for a
// 32-cloud batch, scattered 4-byte index writes).  Here one WARP owns a centre: the cloud is staged in shared memory
// once per CTA, the 32 lanes test 32 consecutive points at a time, a ballot + prefix popcount gives every hit its slot
// (so ascending order is preserved exactly) and the index row [u] is written as contiguous runs; the warp leaves as soon
// as u neighbours are found.  Grouping is a pure gather: lanes walk the (centre, slot) pairs of the index tensor —
// index reads and output writes are fully coalesced, the 4 KB feature rows are read through L1/L2 — and the
// [B,C,M,U] output is written exactly once.
#include "ri_common.cuh"

namespace {

constexpr int kBqThreads = 256;                      // 8 warps = 8 centres in flight per CTA
constexpr int kBqTile = 4096;                        // points staged per pass (64 KB of float4)

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float* __restrict__ centers, const float* __restrict__ points, int n, int m, float r2, int u,
                  int centres_per_cta, int* __restrict__ out)
{
    extern __shared__ float4 spts[];                 // min(n, kBqTile) points
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* Cn = centers + (size_t)b * 3 * m;
    const float* Pt = points + (size_t)b * 3 * n;
    int* O = out + (size_t)b * m * u;
    const int j0 = blockIdx.x * centres_per_cta;
    const int j1 = min(m, j0 + centres_per_cta);
    const int per_warp = (centres_per_cta + kBqThreads / 32 - 1) / (kBqThreads / 32);

    // every warp walks its own centres; the staged tile is shared, so all warps advance tile by tile together
    int cnt[8];                                      // neighbours found so far for this warp's centres (per_warp <= 8)
#pragma unroll
    for (int q = 0; q < 8; ++q) cnt[q] = 0;
    for (int base = 0; base < n; base += kBqTile) {
        const int tile = min(kBqTile, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < tile; t += kBqThreads) {
            const int g = base + t;
            spts[t] = make_float4(Pt[g], Pt[g + n], Pt[g + 2 * (size_t)n], 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + warp * per_warp + q;
            if (q >= per_warp || j >= j1) continue;
            int c = cnt[q];
            if (c >= u) continue;
            const float cx = Cn[j], cy = Cn[j + m], cz = Cn[j + 2 * (size_t)m];
            int* row = O + (size_t)j * u;
            for (int k0 = 0; k0 < tile && c < u; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                if (k < tile) {
                    const float4 p = spts[k];
                    const float dx = __fsub_rn(cx, p.x), dy = __fsub_rn(cy, p.y), dz = __fsub_rn(cz, p.z);
                    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));   // ball_query.cu:36-39
                    hit = d2 < r2 && (double)d2 > 1e-5;                                          // ball_query.cu:40
                }
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (bal == 0u) continue;
                if (c == 0) {                                   // first neighbour fills the whole row (ball_query.cu:41-45)
                    const int first = base + k0 + (__ffs(bal) - 1);
                    for (int v = lane; v < u; v += 32) row[v] = first;
                    __syncwarp();
                }
                const int slot = c + __popc(bal & ((1u << lane) - 1u));
                if (hit && slot < u) row[slot] = base + k;      // ball_query.cu:46 (slots beyond u are never reached)
                c += __popc(bal);
            }
            cnt[q] = c;
        }
    }
    // centres without any neighbour: u zeros
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int j = j0 + warp * per_warp + q;
        if (q >= per_warp || j >= j1 || cnt[q] != 0) continue;
        for (int v = lane; v < u; v += 32) O[(size_t)j * u + v] = 0;
    }
}

// out[b, l, j, k] = feat[b, l, idx[b, j, k]]        (grouping.cu:18-44)
constexpr int kGrpThreads = 256;
constexpr int kGrpChans = 8;
__global__ void __launch_bounds__(kGrpThreads)
grouping_kernel(const float* __restrict__ feat, const int* __restrict__ idx, int c, int n, long long mu,
                float* __restrict__ out)
{
    const int b = blockIdx.z;
    const long long e = (long long)blockIdx.x * kGrpThreads + threadIdx.x;       // (centre, slot) pair, slot fastest
    if (e >= mu) return;
    const int id = __ldg(idx + (size_t)b * mu + e);
    const int l0 = blockIdx.y * kGrpChans, l1 = min(c, l0 + kGrpChans);
    const float* F = feat + (size_t)b * c * n + id;
    float* O = out + (size_t)b * c * mu + e;
    float v[kGrpChans];
#pragma unroll
    for (int q = 0; q < kGrpChans; ++q)
        if (l0 + q < l1) v[q] = __ldg(F + (size_t)(l0 + q) * n);
#pragma unroll
    for (int q = 0; q < kGrpChans; ++q)
        if (l0 + q < l1) O[(size_t)(l0 + q) * mu] = v[q];
}

// grad_x[b, l, idx[b, j, k]] += grad_y[b, l, j, k]   (grouping.cu:58-84: float atomics, as the reference)
__global__ void __launch_bounds__(kGrpThreads)
grouping_grad_kernel(const float* __restrict__ grad_y, const int* __restrict__ idx, int c, int n, long long mu,
                     float* __restrict__ grad_x)
{
    const int b = blockIdx.z;
    const long long e = (long long)blockIdx.x * kGrpThreads + threadIdx.x;
    if (e >= mu) return;
    const int id = __ldg(idx + (size_t)b * mu + e);
    const int l0 = blockIdx.y * kGrpChans, l1 = min(c, l0 + kGrpChans);
    const float* G = grad_y + (size_t)b * c * mu + e;
    float* X = grad_x + (size_t)b * c * n + id;
    for (int l = l0; l < l1; ++l) atomicAdd(X + (size_t)l * n, __ldg(G + (size_t)l * mu));
}

}  // namespace

// ---- C ABI (include/ri_b200.h) ----------------------------------------------------------------------
// ball_query_forward (ball_query/ball_query.cpp:6-30): centers [B,3,M], points [B,3,N] -> neighbours [B,M,U] (int32,
// fully overwritten).  `radius` as the reference takes it; r2 = radius * radius in fp32.
extern "C" int ri_ball_query_f32(const float* centers, const float* points, int B, int N, int M, float radius, int U,
                                 int* neighbors, void* stream)
{
    if (B < 0 || N < 0 || M < 0 || U <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || M == 0) return RI_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        cudaError_t e = cudaMemsetAsync(neighbors, 0, (size_t)B * M * U * sizeof(int), st);
        return e == cudaSuccess ? RI_OK : (int)e;
    }
    const float r2 = radius * radius;                              // ball_query.cpp:24
    const int centres_per_cta = 8 * (kBqThreads / 32);             // 8 centres per warp
    const size_t smem = (size_t)(N < kBqTile ? N : kBqTile) * sizeof(float4);
    RI_KERNEL_SETUP(ball_query_kernel, true, -1);
    dim3 grid((M + centres_per_cta - 1) / centres_per_cta, B);
    ball_query_kernel<<<grid, kBqThreads, smem, st>>>(centers, points, N, M, r2, U, centres_per_cta, neighbors);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Local point-pair features of the shipped models, straight from the neighbour indices.
// Replaces the torch block of PVCNN_classifier.forward, with_local_feat == 'ppf' (/root/reference/PVCNN/models/
// pvcnn_classify.py:252-270) together with the BallQuery module's grouping (PVCNN/modules/ball_query.py:16-35), which
// materialise the grouped coordinates + normals [B,6,U,M] (100 MB at B = 32, U = 128, M = 1024), their expansions and six
// element-wise / reduction passes over them.  Arithmetic kept operation for operation (all fp32, no contraction):
//   rel  = p_nbr - c                                  (ball_query.py:24 — the grouper returns RELATIVE coordinates)
//   d    = c - rel                                    (:262 — so d = 2c - p_nbr: the shipped weights were trained on this)
//   dn   = sqrt((d0^2 + d1^2) + d2^2)                 (:263 torch.norm(dim=1))
//   du   = d / dn                                     (:264 — no epsilon: dn = 0 gives NaN, as in torch)
//   out  = acos(clamp(<n_nbr, du>)), acos(clamp(<n_c, du>)), acos(clamp(<n_nbr, n_c>)), dn      (:265-268)
// with every dot product as torch's mul + sum(dim=1): three rounded products, then (p0 + p1) + p2; acosf is the same
// libdevice function torch calls.  out [B,4,U,M] (the reference's (b, 4, k, m)), written coalesced along M: the index tile
// of 32 centres is transposed through shared memory (indices are [B,M,U], U innermost).
constexpr int kLpThreads = 256;
constexpr int kLpCentres = 32;

__device__ __forceinline__ float lp_dot(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}
__device__ __forceinline__ float lp_acos_clamped(float x)
{
    // torch.clamp propagates NaN; fminf/fmaxf would swallow it
    const float c = (x != x) ? x : fminf(fmaxf(x, -1.0f), 1.0f);
    return acosf(c);
}

__global__ void __launch_bounds__(kLpThreads)
local_ppf_kernel(const float* __restrict__ pc, const float* __restrict__ pn, const float* __restrict__ cc,
                 const float* __restrict__ cn, const int* __restrict__ nbr, int N, int M, int U, float* __restrict__ out)
{
    extern __shared__ int lp_tile[];                               // [kLpCentres][U + 1]
    const int b = blockIdx.y;
    const int m0 = blockIdx.x * kLpCentres;
    const int mc = min(kLpCentres, M - m0);
    const int ld = U + 1;
    const int* I = nbr + ((size_t)b * M + m0) * U;
    for (int e = threadIdx.x; e < mc * U; e += kLpThreads) {       // coalesced along U
        const int ml = e / U, u = e - ml * U;
        lp_tile[ml * ld + u] = I[e];
    }
    __syncthreads();
    const float* PC = pc + (size_t)b * 3 * N;
    const float* PN = pn + (size_t)b * 3 * N;
    const float* CC = cc + (size_t)b * 3 * M;
    const float* CN = cn + (size_t)b * 3 * M;
    const size_t um = (size_t)U * M;
    float* O = out + (size_t)b * 4 * um;
    for (int e = threadIdx.x; e < kLpCentres * U; e += kLpThreads) {
        const int ml = e & (kLpCentres - 1), u = e >> 5;           // lanes walk the centres: coalesced stores
        if (ml >= mc) continue;
        const int m = m0 + ml;
        int j = lp_tile[ml * ld + u];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        const float c0 = CC[m], c1 = CC[m + M], c2 = CC[m + 2 * (size_t)M];
        const float n0 = CN[m], n1 = CN[m + M], n2 = CN[m + 2 * (size_t)M];
        const float q0 = __ldg(PC + j), q1 = __ldg(PC + j + N), q2 = __ldg(PC + j + 2 * (size_t)N);
        const float r0 = __ldg(PN + j), r1 = __ldg(PN + j + N), r2 = __ldg(PN + j + 2 * (size_t)N);
        const float d0 = __fsub_rn(c0, __fsub_rn(q0, c0)), d1 = __fsub_rn(c1, __fsub_rn(q1, c1)), d2 = __fsub_rn(c2, __fsub_rn(q2, c2));
        const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
        const float u0 = __fdiv_rn(d0, dn), u1 = __fdiv_rn(d1, dn), u2 = __fdiv_rn(d2, dn);
        const size_t o = (size_t)u * M + m;
        O[o] = lp_acos_clamped(lp_dot(r0, r1, r2, u0, u1, u2));
        O[um + o] = lp_acos_clamped(lp_dot(n0, n1, n2, u0, u1, u2));
        O[2 * um + o] = lp_acos_clamped(lp_dot(r0, r1, r2, n0, n1, n2));
        O[3 * um + o] = dn;
    }
}

// points_coords / points_normals [B,3,N], centers_coords / centers_normals [B,3,M], neighbors [B,M,U] (ri_ball_query_f32)
// -> out [B,4,U,M]
extern "C" int ri_local_ppf_f32(const float* points_coords, const float* points_normals, const float* centers_coords,
                                const float* centers_normals, const int* neighbors, int B, int N, int M, int U,
                                float* out, void* stream)
{
    if (B < 0 || N < 0 || M < 0 || U < 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || M == 0 || U == 0) return RI_OK;
    if (N == 0) return RI_ERR_BAD_ARG;
    const size_t smem = (size_t)kLpCentres * (U + 1) * sizeof(int);
    if (smem > 200 * 1024) return RI_ERR_UNSUPPORTED;
    RI_KERNEL_SETUP(local_ppf_kernel, true, -1);
    dim3 grid((M + kLpCentres - 1) / kLpCentres, B);
    local_ppf_kernel<<<grid, kLpThreads, smem, (cudaStream_t)stream>>>(points_coords, points_normals, centers_coords,
                                                                      centers_normals, neighbors, N, M, U, out);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// grouping_forward (grouping/grouping.cpp): feat [B,C,N], idx [B,M,U] -> out [B,C,M,U] (fully overwritten)
extern "C" int ri_grouping_f32(const float* feat, const int* idx, int B, int C, int N, int M, int U, float* out, void* stream)
{
    if (B < 0 || C < 0 || N < 0 || M < 0 || U < 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    const long long mu = (long long)M * U;
    if (B == 0 || C == 0 || mu == 0) return RI_OK;
    if (N == 0) return RI_ERR_BAD_ARG;                             // indices into an empty cloud
    const long long bx = (mu + kGrpThreads - 1) / kGrpThreads;
    if (bx > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    dim3 grid((unsigned)bx, (C + kGrpChans - 1) / kGrpChans, B);
    grouping_kernel<<<grid, kGrpThreads, 0, (cudaStream_t)stream>>>(feat, idx, C, N, mu, out);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// grouping_backward: grad_y [B,C,M,U], idx [B,M,U] -> grad_x [B,C,N] (fully overwritten: zeroed, then accumulated)
extern "C" int ri_grouping_backward_f32(const float* grad_y, const int* idx, int B, int C, int N, int M, int U,
                                        float* grad_x, void* stream)
{
    if (B < 0 || C < 0 || N < 0 || M < 0 || U < 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0 || C == 0 || N == 0) return RI_OK;
    cudaError_t e = cudaMemsetAsync(grad_x, 0, (size_t)B * C * N * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    const long long mu = (long long)M * U;
    if (mu == 0) return RI_OK;
    const long long bx = (mu + kGrpThreads - 1) / kGrpThreads;
    if (bx > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
    dim3 grid((unsigned)bx, (C + kGrpChans - 1) / kGrpChans, B);
    grouping_grad_kernel<<<grid, kGrpThreads, 0, st>>>(grad_y, idx, C, N, mu, grad_x);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
