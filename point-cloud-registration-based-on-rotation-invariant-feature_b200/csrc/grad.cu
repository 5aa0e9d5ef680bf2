// grad.cu — backward kernels of the front-end ops, sm_100a.
//
// Replaces  KnnGradKernel                          (/root/reference/PVCNN/modules/functional/src/knn/knn.cu:52-78)
//           avg_voxelize_grad_kernel               (.../voxelization/vox.cu:87-111)
//           spherical_avg_voxelize_grad_kernel     (.../spherical_voxelization/spherical_vox.cu:139-163)
//           trilinear_devoxelize_grad_kernel       (.../interpolate/trilinear_devox.cu:120-163)
//           spherical_trilinear_devoxelize_grad_kernel (.../interpolate/spherical_trilinear_devox.cu:150-194)
// Same math; grids cover (point tiles, channel groups, clouds) instead of one CTA per cloud.  The voxelize
// backward is a pure gather (the reference's atomicAdd onto a zero-filled tensor adds exactly one term per
// element), so it needs neither atomics nor a memset.  The scatter kernels keep float atomics (sum order is not
// part of the contract: 1e-5 relative).
#include "ri_common.cuh"

namespace {

constexpr int kGradThreads = 128;
constexpr int kGradChans = 8;

__global__ void __launch_bounds__(kGradThreads)
knn_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, const float* __restrict__ gdist,
                const int* __restrict__ idx, int c, int n, int m, int k,
                float* __restrict__ grad1, float* __restrict__ grad2)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kGradThreads + threadIdx.x;
    if (i >= n) return;
    const float* A = xyz1 + (size_t)b * c * n;
    const float* Bq = xyz2 + (size_t)b * c * m;
    const float* G = gdist + (size_t)b * k * n;
    const int* I = idx + (size_t)b * k * n;
    float* G1 = grad1 + (size_t)b * c * n;
    float* G2 = grad2 + (size_t)b * c * m;
    for (int q = 0; q < k; ++q) {
        const float g = __fmul_rn(G[i + (size_t)q * n], 2.0f);
        if (g >= 20000.0f) continue;                                   // knn.cu:68
        const int id = I[i + (size_t)q * n];
        for (int p = 0; p < c; ++p) {
            const float t = __fmul_rn(g, __fsub_rn(A[i + (size_t)p * n], Bq[id + (size_t)p * m]));
            atomicAdd(G1 + i + (size_t)p * n, t);
            atomicAdd(G2 + id + (size_t)p * m, -t);
        }
    }
}

__global__ void __launch_bounds__(kGradThreads)
vox_grad_kernel(const float* __restrict__ grad_y, const int* __restrict__ ind, const int* __restrict__ cnt,
                int C, int N, int s, float* __restrict__ grad_x)
{
    const int b = blockIdx.z;
    const int i = blockIdx.x * kGradThreads + threadIdx.x;
    if (i >= N) return;
    const int pos = ind[(size_t)b * N + i];
    const int c0 = blockIdx.y * kGradChans, c1 = min(C, c0 + kGradChans);
    float* GX = grad_x + (size_t)b * C * N + i;
    int n = 0;
    if (pos >= 0 && pos < s) n = cnt[(size_t)b * s + pos];
    if (n <= 0) {
        for (int c = c0; c < c1; ++c) GX[(size_t)c * N] = 0.f;
        return;
    }
    const float inv = __fdiv_rn(1.0f, (float)n);
    const float* GY = grad_y + (size_t)b * C * s + pos;
    for (int c = c0; c < c1; ++c)
        GX[(size_t)c * N] = __fadd_rn(0.f, __fmul_rn(__ldg(GY + (size_t)c * s), inv));
}

__global__ void __launch_bounds__(kGradThreads)
devox_grad_kernel(const float* __restrict__ grad_y, const int* __restrict__ inds, const float* __restrict__ wgts,
                  int C, int N, int s, int skip_undefined, float* __restrict__ grad_x)
{
    const int b = blockIdx.z;
    const int i = blockIdx.x * kGradThreads + threadIdx.x;
    if (i >= N) return;
    const int* I = inds + (size_t)b * 8 * N + i;
    const float* Wt = wgts + (size_t)b * 8 * N + i;
    int id[8]; float w[8];
    id[0] = I[0];
    if (skip_undefined && id[0] == -1) return;
#pragma unroll
    for (int q = 1; q < 8; ++q) id[q] = I[(size_t)q * N];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = Wt[(size_t)q * N];
    const int c0 = blockIdx.y * kGradChans, c1 = min(C, c0 + kGradChans);
    for (int c = c0; c < c1; ++c) {
        const float g = grad_y[((size_t)b * C + c) * N + i];
        float* GX = grad_x + ((size_t)b * C + c) * s;
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(GX + id[q], __fmul_rn(w[q], g));
    }
}

}  // namespace

extern "C" int ri_knn_backward_f32(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                                   const int* idx1, const int* idx2, int B, int c, int n, int m, int k,
                                   float* gradxyz1, float* gradxyz2, void* stream)
{
    if (B < 0 || c <= 0 || n < 0 || m < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(gradxyz1, 0, (size_t)B * c * n * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(gradxyz2, 0, (size_t)B * c * m * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (B == 0) return RI_OK;
    if (n > 0) {
        dim3 g((n + kGradThreads - 1) / kGradThreads, B);
        knn_grad_kernel<<<g, kGradThreads, 0, st>>>(xyz1, xyz2, graddist1, idx1, c, n, m, k, gradxyz1, gradxyz2);
        RI_LAUNCH_CHECK();
    }
    if (m > 0) {
        dim3 g((m + kGradThreads - 1) / kGradThreads, B);
        knn_grad_kernel<<<g, kGradThreads, 0, st>>>(xyz2, xyz1, graddist2, idx2, c, m, n, k, gradxyz2, gradxyz1);
        RI_LAUNCH_CHECK();
    }
    return RI_OK;
}

extern "C" int ri_voxelize_backward_f32(const float* grad_y, const int* ind, const int* cnt, int B, int C, int N, int s,
                                        float* grad_x, void* stream)
{
    if (B < 0 || C < 0 || N < 0 || s <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || C == 0 || N == 0) return RI_OK;
    dim3 g((N + kGradThreads - 1) / kGradThreads, (C + kGradChans - 1) / kGradChans, B);
    vox_grad_kernel<<<g, kGradThreads, 0, (cudaStream_t)stream>>>(grad_y, ind, cnt, C, N, s, grad_x);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

extern "C" int ri_devox_backward_f32(const float* grad_y, const int* inds, const float* wgts, int B, int C, int N, int s,
                                     int skip_undefined, float* grad_x, void* stream)
{
    if (B < 0 || C < 0 || N < 0 || s <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_x, 0, (size_t)B * C * s * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (B == 0 || C == 0 || N == 0) return RI_OK;
    dim3 g((N + kGradThreads - 1) / kGradThreads, (C + kGradChans - 1) / kGradChans, B);
    devox_grad_kernel<<<g, kGradThreads, 0, st>>>(grad_y, inds, wgts, C, N, s, skip_undefined, grad_x);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
