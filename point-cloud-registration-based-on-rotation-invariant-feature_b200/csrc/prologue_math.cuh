// prologue_math.cuh — arithmetic of the voxelization modules' coordinate prologue, shared by prologue.cu and the fused
// front kernel of voxelize.cu (Voxelization.forward /root/reference/PVCNN/modules/voxelization.py:16-35,
// Spherical_Voxelization.forward /root/reference/PVCNN/modules/spherical_vox.py:14-23).
#pragma once
#include "ri_common.cuh"

constexpr int kProThreads = 512;

__device__ __forceinline__ float radius3(float x, float y, float z, int mode)
{
    float q;
    switch (mode) {
        case 0: q = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))); break;
        case 1: q = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)); break;
        case 2: q = __fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))); break;
        case 3: q = __fmaf_rn(z, z, __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); break;
        default: q = __fmaf_rn(x, x, __fmaf_rn(y, y, __fmul_rn(z, z))); break;
    }
    return __fsqrt_rn(q);
}

__device__ __forceinline__ float block_max(float v, float* sred)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = sred[0];
    for (int w = 1; w < kProThreads / 32; ++w) m = fmaxf(m, sred[w]);
    __syncthreads();
    return m;
}


// One coordinate after centring: shape 0 = cube normalize=False, 1 = cube normalize=True, 2 = spherical.
// Returns the value stored in norm_coords; for the cube shapes *vox gets round-half-even of it (torch.round).
__device__ __forceinline__ float ri_prologue_coord(float centred, int shape, float denom, float rf, float hi, int* vox)
{
    float v = centred;
    if (shape == 2) return __fdiv_rn(v, denom);
    if (shape == 0) v = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);          // (nc + 1) / 2.0
    else v = __fadd_rn(__fdiv_rn(v, denom), 0.5f);
    v = fminf(fmaxf(__fmul_rn(v, rf), 0.0f), hi);                        // clamp(nc * r, 0, r - 1)
    *vox = __float2int_rn(rintf(v));
    return v;
}
