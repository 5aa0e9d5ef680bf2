// prologue.cu — the coordinate prologue of the voxelization modules as ONE kernel, sm_100a.
//
// Replaces the chain of small torch kernels in
//   Voxelization.forward            (/root/reference/PVCNN/modules/voxelization.py:16-35)
//   Spherical_Voxelization.forward  (/root/reference/PVCNN/modules/spherical_vox.py:14-23)
// i.e. centre -> scale -> (clamp, round | divide by the largest radius).  The per-cloud mean over the N points is
// NOT recomputed here: it is passed in (torch's own `coords.mean(2)`), because the order of that 1024-term sum is
// what defines the low bits of everything downstream.  Everything after the mean is elementwise IEEE arithmetic
// (sub, add, mul, div, min/max, rint) or an exact max-reduction, which this kernel reproduces bit for bit; the one
// 3-term sum (the point radius) follows the association torch's norm kernel uses (norm_mode, pinned by
// tests/test_parity_gpu.py::test_prologue_bit_identical_to_torch).
//
// One CTA per cloud; input is the interleaved [B,6,N] (xyz | normal) batch, outputs are the contiguous planes the
// other kernels take: xyz [B,3,N], normals [B,3,N], norm_coords [B,3,N] and, for the cube variant, int voxel coords.
#include "ri_common.cuh"
#include "prologue_math.cuh"

namespace {

// shape: 0 = cube, normalize=False   ((x - mean + 1) / 2 * r, clamp, round)
//        1 = cube, normalize=True    ((x - mean) / (2 * maxradius + eps) + 0.5, * r, clamp, round)
//        2 = spherical               ((x - mean) / (maxradius + 1e-20))
__global__ void __launch_bounds__(kProThreads)
prologue_kernel(const float* __restrict__ points, int pstride, const float* __restrict__ mean, int N, int r,
                int shape, float eps, int norm_mode,
                float* __restrict__ xyz, float* __restrict__ normals,
                float* __restrict__ norm_coords, int* __restrict__ vox_coords)
{
    __shared__ float sred[kProThreads / 32];
    const int b = blockIdx.x;
    const float* P = points + (size_t)b * pstride * N;
    const float mx = mean[b * 3 + 0], my = mean[b * 3 + 1], mz = mean[b * 3 + 2];
    const size_t o3 = (size_t)b * 3 * N;
    float denom = 1.0f;
    if (shape != 0) {
        float m = 0.f;                                        // radii are >= 0
        for (int i = threadIdx.x; i < N; i += kProThreads)
            m = fmaxf(m, radius3(__fsub_rn(P[i], mx), __fsub_rn(P[i + N], my), __fsub_rn(P[i + 2 * (size_t)N], mz), norm_mode));
        m = block_max(m, sred);
        denom = (shape == 1) ? __fadd_rn(__fmul_rn(m, 2.0f), eps) : __fadd_rn(m, 1e-20f);
    }
    const float rf = (float)r, hi = (float)(r - 1);
    for (int i = threadIdx.x; i < N; i += kProThreads) {
        float c[3];
        c[0] = P[i]; c[1] = P[i + N]; c[2] = P[i + 2 * (size_t)N];
        if (xyz != nullptr) { xyz[o3 + i] = c[0]; xyz[o3 + N + i] = c[1]; xyz[o3 + 2 * (size_t)N + i] = c[2]; }
        if (normals != nullptr && pstride >= 6) {
            normals[o3 + i] = P[i + 3 * (size_t)N];
            normals[o3 + N + i] = P[i + 4 * (size_t)N];
            normals[o3 + 2 * (size_t)N + i] = P[i + 5 * (size_t)N];
        }
        c[0] = __fsub_rn(c[0], mx); c[1] = __fsub_rn(c[1], my); c[2] = __fsub_rn(c[2], mz);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = c[a];
            if (shape == 2) {
                v = __fdiv_rn(v, denom);
            } else {
                if (shape == 0) v = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);          // (nc + 1) / 2.0
                else v = __fadd_rn(__fdiv_rn(v, denom), 0.5f);
                v = fminf(fmaxf(__fmul_rn(v, rf), 0.0f), hi);                        // clamp(nc * r, 0, r - 1)
                vox_coords[o3 + (size_t)a * N + i] = __float2int_rn(rintf(v));       // torch.round: half to even
            }
            norm_coords[o3 + (size_t)a * N + i] = v;
        }
    }
}

}  // namespace

// points [B,pstride,N] with pstride in {3, 6}; mean [B,3] = per-cloud mean of the xyz planes (torch's).
// xyz / normals may be null (no split wanted); vox_coords is required for shape 0/1 and ignored for shape 2.
extern "C" int ri_vox_prologue_f32(const float* points, int pstride, const float* mean, int B, int N, int r,
                                   int shape, float eps, int norm_mode,
                                   float* xyz, float* normals, float* norm_coords, int* vox_coords, void* stream)
{
    if (B < 0 || N < 0 || r <= 0 || (pstride != 3 && pstride != 6) || shape < 0 || shape > 2) return RI_ERR_BAD_ARG;
    if (norm_coords == nullptr || (shape != 2 && vox_coords == nullptr)) return RI_ERR_BAD_ARG;
    if (B == 0 || N == 0) return RI_OK;
    RI_KERNEL_SETUP(prologue_kernel, false, ri_step_carveout_percent());
    prologue_kernel<<<B, kProThreads, 0, (cudaStream_t)stream>>>(points, pstride, mean, N, r, shape, eps, norm_mode,
                                                                 xyz, normals, norm_coords, vox_coords);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// De-interleave a [B, 6, N] (xyz | normal) batch into the contiguous xyz [B,3,N] and normals [B,3,N] planes the k-NN and
// PPF kernels take — the `.contiguous()` copies of the slices inputs[:, :3, :] / inputs[:, 3:, :] that the reference's
// wrappers make (functional/knn.py:11-12, functional/ppf.py:16-19) — and, optionally, the point-major packing
// packed [B,N,8] = (x, y, z, nx, ny, nz, 0, 0) that ri_ppf_gather_packed_f32 gathers from; one launch.
namespace {
__global__ void __launch_bounds__(256)
split6_kernel(const float* __restrict__ points, int N, float* __restrict__ xyz, float* __restrict__ normals,
              float4* __restrict__ packed)
{
    const size_t b = blockIdx.y;
    const float* P = points + b * 6 * N;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        float v[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = P[(size_t)a * N + i];
        if (xyz != nullptr) {
#pragma unroll
            for (int a = 0; a < 3; ++a) xyz[b * 3 * N + (size_t)a * N + i] = v[a];
        }
        if (normals != nullptr) {
#pragma unroll
            for (int a = 0; a < 3; ++a) normals[b * 3 * N + (size_t)a * N + i] = v[3 + a];
        }
        if (packed != nullptr) {
            packed[(b * N + i) * 2] = make_float4(v[0], v[1], v[2], v[3]);
            packed[(b * N + i) * 2 + 1] = make_float4(v[4], v[5], 0.f, 0.f);
        }
    }
}
}  // namespace

extern "C" int ri_split_xyz_normals_f32(const float* points, int B, int N, float* xyz, float* normals, float* packed,
                                        void* stream)
{
    if (B < 0 || N < 0 || B > 65535 || ((uintptr_t)packed & 15) != 0) return RI_ERR_BAD_ARG;
    if (B == 0 || N == 0) return RI_OK;
    RI_KERNEL_SETUP(split6_kernel, false, ri_step_carveout_percent());
    dim3 grid((unsigned)((N + 255) / 256 > 16 ? 16 : (N + 255) / 256), B);
    split6_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, N, xyz, normals, reinterpret_cast<float4*>(packed));
    RI_LAUNCH_CHECK();
    return RI_OK;
}
