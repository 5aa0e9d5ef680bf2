// lrf.cu — the rotation-invariant 'change_coords' preprocessing (global local-reference-frame) as ONE kernel, sm_100a.
// SURVEY.md §8 row f4.
//
// Replaces the Python double loop of PVCNN_classifier.forward, rot_invariant_preprocess == 'change_coords'
// (/root/reference/PVCNN/models/pvcnn_classify.py:153-184), which walks the batch and, per cloud, the points in order of
// decreasing radius with a `.norm()` host synchronisation per step:
//   nc      = coords - mean(coords)                                          (:154)
//   rank    = argsort(|nc|, descending)                                       (:155)
//   base_x  = nc[rank[0]] / |nc[rank[0]]|                                     (:159-161)
//   base_y  = the first nc[rank[j]], j >= 1, with |.| >= 1e-5, normalised, whose lamda = <base_x, base_y> lies in (-0.9, 0.9)  (:162-169)
//   base_x -= base_y * <base_x, base_y>;  base_x /= |base_x|                  (:175-177)   (Gram-Schmidt, y kept)
//   base_z  = base_x x base_y;  base_z /= |base_z|                            (:179-180)
//   new     = (base_x . nc, base_y . nc, base_z . nc)                         (:181-184)
// The per-cloud mean is passed in (torch's own reduction, as for the voxelizer prologue) so that the radius ranking sees
// the same bits.  "First in descending-radius order that passes the test" is a max-reduction over the passing points —
// no sort: two block-wide arg-max passes with packed (radius bits, ~index) keys (the lowest index wins equal radii;
// torch.argsort leaves that case unspecified).  A cloud without an admissible base_y, or whose largest radius is below 1e-5,
// trips an `assert` in the reference; here ok[b] = 0 and its output is zeros.
#include "ri_common.cuh"
#include "prologue_math.cuh"

namespace {

constexpr int kLrfThreads = 512;

__device__ __forceinline__ unsigned long long lrf_block_max(unsigned long long v, unsigned long long* sred)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long m = sred[0];
    for (int w = 1; w < kLrfThreads / 32; ++w) m = sred[w] > m ? sred[w] : m;
    __syncthreads();
    return m;
}

__device__ __forceinline__ unsigned long long lrf_key(float radius, int i)
{
    return ((unsigned long long)__float_as_uint(radius) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);   // radius >= 0
}

__device__ __forceinline__ float sum3(float a, float b, float c) { return __fadd_rn(__fadd_rn(a, b), c); }

__global__ void __launch_bounds__(kLrfThreads)
lrf_kernel(const float* __restrict__ coords, int cstride, const float* __restrict__ mean, int N, int norm_mode,
           float* __restrict__ out, float* __restrict__ bases, int* __restrict__ ok)
{
    __shared__ unsigned long long sred[kLrfThreads / 32];
    const int b = blockIdx.x;
    const float* X = coords + (size_t)b * cstride * N;
    const float mx = mean[3 * b], my = mean[3 * b + 1], mz = mean[3 * b + 2];
    float* O = out + (size_t)b * 3 * N;

    // pass 1: the farthest point
    unsigned long long k = 0ull;
    for (int i = threadIdx.x; i < N; i += kLrfThreads) {
        const float r = radius3(__fsub_rn(X[i], mx), __fsub_rn(X[i + N], my), __fsub_rn(X[i + 2 * (size_t)N], mz), norm_mode);
        const unsigned long long ki = lrf_key(r, i);
        k = ki > k ? ki : k;
    }
    k = lrf_block_max(k, sred);
    const int i0 = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
    float bx[3] = {__fsub_rn(X[i0], mx), __fsub_rn(X[i0 + N], my), __fsub_rn(X[i0 + 2 * (size_t)N], mz)};
    const float n0 = radius3(bx[0], bx[1], bx[2], 1);
    bool good = N >= 2 && n0 > 1e-5f;                                                 // :160 assert
    bx[0] = __fdiv_rn(bx[0], n0); bx[1] = __fdiv_rn(bx[1], n0); bx[2] = __fdiv_rn(bx[2], n0);

    // pass 2: the farthest of the other points whose direction is not (anti)parallel to base_x
    k = 0ull;
    for (int i = threadIdx.x; i < N; i += kLrfThreads) {
        if (i == i0) continue;
        const float cx = __fsub_rn(X[i], mx), cy = __fsub_rn(X[i + N], my), cz = __fsub_rn(X[i + 2 * (size_t)N], mz);
        const float r = radius3(cx, cy, cz, norm_mode);                               // the ranking radius (:155)
        const float rn = radius3(cx, cy, cz, 1);                                      // base_y.norm() (:164, :166)
        if (rn < 1e-5f) continue;
        const float lam = sum3(__fmul_rn(bx[0], __fdiv_rn(cx, rn)), __fmul_rn(bx[1], __fdiv_rn(cy, rn)),
                               __fmul_rn(bx[2], __fdiv_rn(cz, rn)));                  // :167
        if (lam < 0.9f && lam > -0.9f) {
            const unsigned long long ki = lrf_key(r, i) | (1ull << 63);               // bit 63: "a candidate exists" (radius >= 0)
            k = ki > k ? ki : k;
        }
    }
    k = lrf_block_max(k, sred);
    good = good && (k >> 63) != 0ull;
    float by[3] = {0.f, 0.f, 0.f}, bz[3] = {0.f, 0.f, 0.f};
    if (good) {
        const int i1 = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
        by[0] = __fsub_rn(X[i1], mx); by[1] = __fsub_rn(X[i1 + N], my); by[2] = __fsub_rn(X[i1 + 2 * (size_t)N], mz);
        const float n1 = radius3(by[0], by[1], by[2], 1);
        by[0] = __fdiv_rn(by[0], n1); by[1] = __fdiv_rn(by[1], n1); by[2] = __fdiv_rn(by[2], n1);
        const float d = sum3(__fmul_rn(bx[0], by[0]), __fmul_rn(bx[1], by[1]), __fmul_rn(bx[2], by[2]));    // :175 bmm
        bx[0] = __fsub_rn(bx[0], __fmul_rn(by[0], d)); bx[1] = __fsub_rn(bx[1], __fmul_rn(by[1], d));
        bx[2] = __fsub_rn(bx[2], __fmul_rn(by[2], d));
        const float nx = radius3(bx[0], bx[1], bx[2], norm_mode);                     // :177 norm(dim=1)
        good = !(nx < 1e-5f);                                                          // :176 assert
        bx[0] = __fdiv_rn(bx[0], nx); bx[1] = __fdiv_rn(bx[1], nx); bx[2] = __fdiv_rn(bx[2], nx);
        bz[0] = __fsub_rn(__fmul_rn(bx[1], by[2]), __fmul_rn(bx[2], by[1]));          // :179 cross
        bz[1] = __fsub_rn(__fmul_rn(bx[2], by[0]), __fmul_rn(bx[0], by[2]));
        bz[2] = __fsub_rn(__fmul_rn(bx[0], by[1]), __fmul_rn(bx[1], by[0]));
        const float nz = radius3(bz[0], bz[1], bz[2], norm_mode);
        bz[0] = __fdiv_rn(bz[0], nz); bz[1] = __fdiv_rn(bz[1], nz); bz[2] = __fdiv_rn(bz[2], nz);
    }
    if (threadIdx.x == 0) {
        ok[b] = good ? 1 : 0;
        if (bases != nullptr) {
            float* Bs = bases + (size_t)b * 9;
#pragma unroll
            for (int a = 0; a < 3; ++a) { Bs[a] = good ? bx[a] : 0.f; Bs[3 + a] = good ? by[a] : 0.f; Bs[6 + a] = good ? bz[a] : 0.f; }
        }
    }
    for (int i = threadIdx.x; i < N; i += kLrfThreads) {
        const float cx = __fsub_rn(X[i], mx), cy = __fsub_rn(X[i + N], my), cz = __fsub_rn(X[i + 2 * (size_t)N], mz);
        // 1 x 3 by 3 x n products (:181-183)
        O[i] = good ? __fmaf_rn(bx[2], cz, __fmaf_rn(bx[1], cy, __fmul_rn(bx[0], cx))) : 0.f;
        O[i + N] = good ? __fmaf_rn(by[2], cz, __fmaf_rn(by[1], cy, __fmul_rn(by[0], cx))) : 0.f;
        O[i + 2 * (size_t)N] = good ? __fmaf_rn(bz[2], cz, __fmaf_rn(bz[1], cy, __fmul_rn(bz[0], cx))) : 0.f;
    }
}

}  // namespace

// coords [B,cstride,N] (cstride 3, or 6 for the interleaved xyz | normal batch; the first 3 planes are used), mean [B,3] =
// torch's coords.mean(2)  ->  out [B,3,N] the coordinates in the cloud's own frame, bases [B,3,3] (rows base_x, base_y,
// base_z; may be NULL), ok [B] (0 where the reference would fail its asserts).
extern "C" int ri_lrf_change_coords_f32(const float* coords, int cstride, const float* mean, int B, int N, int norm_mode,
                                        float* out, float* bases, int* ok, void* stream)
{
    if (B < 0 || N < 0 || (cstride != 3 && cstride != 6) || out == nullptr || ok == nullptr) return RI_ERR_BAD_ARG;
    if (B == 0) return RI_OK;
    if (N == 0) { cudaError_t e = cudaMemsetAsync(ok, 0, sizeof(int) * (size_t)B, (cudaStream_t)stream); return e == cudaSuccess ? RI_OK : (int)e; }
    lrf_kernel<<<B, kLrfThreads, 0, (cudaStream_t)stream>>>(coords, cstride, mean, N, norm_mode, out, bases, ok);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
