// devox.cu — cube and spherical trilinear devoxelization + the DGCNN voxel-neighbour edge gather, sm_100a.
//
// Replaces  trilinear_devoxelize_kernel            (/root/reference/PVCNN/modules/functional/src/interpolate/
//                                                   trilinear_devox.cu:22-106, host trilinear_devox.cpp:18-55)
//           spherical_trilinear_devoxelize_kernel  (.../interpolate/spherical_trilinear_devox.cu:23-136,
//                                                   host spherical_trilinear_devox.cpp:19-56)
//           the torch gather/sub/mask/cat block of PVConv.forward (/root/reference/PVCNN/modules/pvconv.py:68-90).
//
// The spherical variant keeps the reference's index quirks on purpose (integer `grid_gama / r` == 0, radian
// values truncated to ints and used as grid rows/columns, residuals not normalised by the cell size): they are
// what the shipped weights were trained against.
//
// Launch shape: grid = (point tiles, channel groups, clouds) — a 1024-point, 64-channel cloud becomes 64 CTAs
// instead of the reference's single 512-thread CTA, and each thread keeps 8 channels x 8 corners = 64 independent
// gathers in flight.  Lanes of a warp are consecutive points, so outs/inds/wgts stores are fully coalesced.
// The 8-term sum uses the reference's contraction order (001 mul, then fma 000,010,011,100,101,110,111), so outs
// is bit-identical given identical inputs.
#include "ri_common.cuh"

namespace {

constexpr int kDevoxThreads = 128;
constexpr int kDevoxChans = 8;

__device__ __forceinline__ float devox_sum(const float* __restrict__ f, const int (&id)[8], const float (&w)[8])
{
    float acc = __fmul_rn(w[1], __ldg(f + id[1]));
    acc = __fmaf_rn(w[0], __ldg(f + id[0]), acc);
    acc = __fmaf_rn(w[2], __ldg(f + id[2]), acc);
    acc = __fmaf_rn(w[3], __ldg(f + id[3]), acc);
    acc = __fmaf_rn(w[4], __ldg(f + id[4]), acc);
    acc = __fmaf_rn(w[5], __ldg(f + id[5]), acc);
    acc = __fmaf_rn(w[6], __ldg(f + id[6]), acc);
    acc = __fmaf_rn(w[7], __ldg(f + id[7]), acc);
    return acc;
}

// SPH == false: coords are grid-unit coordinates in [0, r-1] (Voxelization.forward's norm_coords).
// SPH == true : coords are the normalised Cartesian coords, g_inds the spherical cell of each point.
template <bool SPH>
__global__ void __launch_bounds__(kDevoxThreads)
devox_kernel(const float* __restrict__ coords, const float* __restrict__ feat, const int* __restrict__ g_inds,
             int C, int N, int r, float* __restrict__ outs, int* __restrict__ inds, float* __restrict__ wgts)
{
    // Clouds are walked in DESCENDING order: when the grid was produced just before by a kernel that wrote clouds
    // in ascending order (the voxelizer, a Conv3d), the highest-numbered clouds are the ones still resident in L2.
    const int b = (int)gridDim.z - 1 - (int)blockIdx.z;
    const int i = blockIdx.x * kDevoxThreads + threadIdx.x;
    if (i >= N) return;
    const int r2 = r * r;
    const size_t s = (size_t)r2 * r;
    const float* X = coords + (size_t)b * 3 * N;
    const bool writer = blockIdx.y == 0;         // channel group 0 also emits inds / wgts
    int id[8]; float w[8];
    bool defined = true;
    int first_ind = 0;                           // what inds[0,i] holds for an undefined point

    if (SPH) {
        const int pos = g_inds[(size_t)b * N + i];
        float g = 0.f, a = 0.f, be = 0.f;
        if (pos == -1) { defined = false; first_ind = -1; }                         // :42-47
        else if (!ri_sph_coords(X[i], X[i + N], X[i + 2 * (size_t)N], r, g, a, be)) defined = false;   // :54,:59
        if (defined) {
            const int gg = pos / r2;
            const int ga = (pos - gg * r2) / r;
            const int gb = pos - gg * r2 - ga * r;
            const float g_lo = (float)(gg / r);                                                         // :71
            const float a_lo = __double2float_rn(__ddiv_rn(__dmul_rn(__dmul_rn(RI_PI, 2.0), (double)ga), (double)r));  // :72
            const float b_lo = __double2float_rn(__ddiv_rn(__dmul_rn(RI_PI, (double)gb), (double)r));   // :73
            ri_corners(__fsub_rn(g, g_lo), __fsub_rn(a, a_lo), __fsub_rn(be, b_lo),
                       (int)g_lo, (int)a_lo, (int)b_lo, r, r2, id, w);
        }
    } else {
        const float x = X[i], y = X[i + N], z = X[i + 2 * (size_t)N];
        const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
        ri_corners(__fsub_rn(x, xl), __fsub_rn(y, yl), __fsub_rn(z, zl), (int)xl, (int)yl, (int)zl, r, r2, id, w);
    }

    if (writer) {
        int* I = inds + (size_t)b * 8 * N + i;
        float* Wt = wgts + (size_t)b * 8 * N + i;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            I[(size_t)q * N] = defined ? id[q] : (q == 0 ? first_ind : 0);
            Wt[(size_t)q * N] = defined ? w[q] : 0.f;
        }
    }
    const int c0 = blockIdx.y * kDevoxChans;
    const int c1 = min(C, c0 + kDevoxChans);
    float* O = outs + (size_t)b * C * N + i;
    if (!defined) {
        for (int c = c0; c < c1; ++c) O[(size_t)c * N] = 0.f;       // the reference leaves its zero-fill
        return;
    }
    const float* F = feat + (size_t)b * C * s;
    if (c1 - c0 == kDevoxChans) {
        float v[kDevoxChans];
#pragma unroll
        for (int u = 0; u < kDevoxChans; ++u) v[u] = devox_sum(F + (size_t)(c0 + u) * s, id, w);
#pragma unroll
        for (int u = 0; u < kDevoxChans; ++u) O[(size_t)(c0 + u) * N] = v[u];
    } else {
        for (int c = c0; c < c1; ++c) O[(size_t)c * N] = devox_sum(F + (size_t)c * s, id, w);
    }
}

// out[b, c, i]     = inds[b,i] == -1 ? 0 : feat[b,c,i] - avg[b,c,inds[b,i]]
// out[b, C + c, i] = feat[b,c,i]                                                    (pvconv.py:68-90)
constexpr int kEdgeThreads = 128;
constexpr int kEdgeChans = 8;
__global__ void __launch_bounds__(kEdgeThreads)
edge_gather_kernel(const float* __restrict__ avg, const float* __restrict__ feat, const int* __restrict__ inds,
                   int C, int N, int s, float* __restrict__ out)
{
    const int b = blockIdx.z;
    const int i = blockIdx.x * kEdgeThreads + threadIdx.x;
    if (i >= N) return;
    const int id = inds[(size_t)b * N + i];
    const bool undef = id == -1;
    const int idc = undef ? 0 : id;
    const int c0 = blockIdx.y * kEdgeChans, c1 = min(C, c0 + kEdgeChans);
    const float* A = avg + (size_t)b * C * s + idc;
    const float* F = feat + (size_t)b * C * N + i;
    float* O = out + (size_t)b * 2 * C * N + i;
    float f[kEdgeChans], a[kEdgeChans];
#pragma unroll
    for (int u = 0; u < kEdgeChans; ++u) {
        const int c = c0 + u;
        if (c < c1) { f[u] = F[(size_t)c * N]; a[u] = __ldg(A + (size_t)c * s); }
    }
#pragma unroll
    for (int u = 0; u < kEdgeChans; ++u) {
        const int c = c0 + u;
        if (c < c1) {
            O[(size_t)c * N] = undef ? 0.f : __fsub_rn(f[u], a[u]);
            O[(size_t)(C + c) * N] = f[u];
        }
    }
}

template <bool SPH>
int devox_impl(const float* coords, const float* feat, const int* g_inds, int B, int C, int N, int r,
               float* outs, int* inds, float* wgts, cudaStream_t st)
{
    if (B < 0 || C < 0 || N < 0 || r <= 0 || r > 1024) return RI_ERR_BAD_ARG;
    if ((long long)r * r * r > 0x7fffffffLL || B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || N == 0) return RI_OK;
    const int groups = C > 0 ? (C + kDevoxChans - 1) / kDevoxChans : 1;   // one group still writes inds/wgts
    dim3 grid((N + kDevoxThreads - 1) / kDevoxThreads, groups, B);
    devox_kernel<SPH><<<grid, kDevoxThreads, 0, st>>>(coords, feat, g_inds, C, N, r, outs, inds, wgts);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

}  // namespace

extern "C" int ri_trilinear_devox_f32(const float* coords, const float* feat, int B, int C, int N, int r,
                                      float* outs, int* inds, float* wgts, void* stream)
{
    return devox_impl<false>(coords, feat, nullptr, B, C, N, r, outs, inds, wgts, (cudaStream_t)stream);
}

extern "C" int ri_sph_trilinear_devox_f32(const float* coords, const float* feat, const int* g_inds,
                                          int B, int C, int N, int r, float* outs, int* inds, float* wgts, void* stream)
{
    if (g_inds == nullptr) return RI_ERR_BAD_ARG;
    return devox_impl<true>(coords, feat, g_inds, B, C, N, r, outs, inds, wgts, (cudaStream_t)stream);
}

extern "C" int ri_voxel_edge_gather_f32(const float* avg, const float* feat, const int* inds,
                                        int B, int C, int N, int s, float* out, void* stream)
{
    if (B < 0 || C < 0 || N < 0 || s <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || N == 0 || C == 0) return RI_OK;
    dim3 grid((N + kEdgeThreads - 1) / kEdgeThreads, (C + kEdgeChans - 1) / kEdgeChans, B);
    edge_gather_kernel<<<grid, kEdgeThreads, 0, (cudaStream_t)stream>>>(avg, feat, inds, C, N, s, out);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
