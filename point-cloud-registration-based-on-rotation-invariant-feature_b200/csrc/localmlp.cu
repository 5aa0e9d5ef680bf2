// localmlp.cu — the local-feature branch of the shipped models in ONE kernel: neighbour indices -> local point-pair features
// -> SharedMLP(4 -> 32 -> 64) (eval mode, BatchNorm folded into the 1x1 convolutions) -> max over the neighbours.  sm_100a.
//
// Replaces, for with_local_feat == 'ppf' (/root/reference/PVCNN/models/pvcnn_classify.py:61-67, 252-271):
//     neighbor_coords_normals = self.grouper(coords, center_coords, normals)        # BallQuery + grouping  [b, 6, k, n]
//     ... local_ppf = cat(nr_d, ni_d, nr_ni, d_norm)                                 # [b, 4, k, n]
//     local_features = self.fuser(local_ppf).max(dim=2).values                       # [b, 64, n]
// where fuser = SharedMLP(4, [32, 64], dim=2) = Conv2d(4,32,1) BN ReLU Conv2d(32,64,1) BN ReLU (modules/shared_mlp.py:6-31).
// The reference materialises local_ppf (67 MB at b = 32, k = 128, n = 1024), the 32-channel activation (537 MB) and the
// 64-channel activation (1.07 GB) before reducing over k; here nothing between the indices and the [b, 64, n] result
// leaves the SM.
//
// One CTA of 128 threads works on one centre at a time; thread j is neighbour j (k = 128 = the M of one tcgen05 MMA):
//   * point-pair features of (centre, neighbour j): the arithmetic of local_ppf_kernel (ballquery.cu), operation for
//     operation;
//   * layer 1 (4 -> 32, 6 % of the flops) on the CUDA cores in fp32, bias + ReLU, result split hi / lo
//     (hi = the value with its low 13 mantissa bits cleared = what a tf32 operand keeps, lo = value - hi, exact) and
//     written to shared memory in the UMMA canonical K-major layout as the A operand [128 x 32];
//   * layer 2 (32 -> 64) on the tensor cores: 3xTF32 split product  A W^T ~ lo.hi + hi.lo + hi.hi  (12 tcgen05.mma
//     kind::tf32 M128 N64 K8, fp32 accumulation in TMEM, 64 columns), W (hi / lo) staged once per CTA;
//   * epilogue: tcgen05.ld of the thread's row, bias + ReLU, max over the 128 rows (redux.sync on the bit patterns — the
//     values are >= +0 —, then across the four warps through shared memory), 64 floats stored per centre.
// Several CTAs per SM (48 KB of shared memory, 64 TMEM columns each) overlap each other's latency phases.
#include "ri_common.cuh"

namespace {

constexpr int kU = 128;                 // neighbours per centre = MMA M
constexpr int kC1 = 32, kC2 = 64;       // hidden / output channels
constexpr int kThreads = 128;
constexpr unsigned kLBO = 128, kSBO = 512;                // K-major, no swizzle: k-core stride, 8-row group stride
constexpr int kAPlane = kU * kC1 * 4;                     // 16 KB: [2 chunks of 16 k][128 rows x 64 B]
constexpr int kBPlane = kC2 * kC1 * 4;                    // 8 KB
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kC2 >> 3) << 17) | ((uint32_t)(kU >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, no swizzle, descriptor version 1 (Blackwell): see matcher.cu::smem_desc
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of element (row, k) of an [R x 32] K-major operand: two 16-k chunks of R x 64 B, 8-row groups of 512 B,
// 16-byte k-cores 128 B apart
__device__ __forceinline__ uint32_t op_offset(int rows, int row, int k)
{
    return (uint32_t)((k >> 4) * rows * 64 + (row >> 3) * kSBO + ((k & 15) >> 2) * kLBO + (row & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ float lp_dot(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}
__device__ __forceinline__ float lp_acos_clamped(float x)
{
    const float c = (x != x) ? x : fminf(fmaxf(x, -1.0f), 1.0f);      // torch.clamp propagates NaN
    return acosf(c);
}
// ReLU that keeps NaN (torch's does) and never returns -0 (the maximum below is taken on the bit patterns)
__device__ __forceinline__ float relu_pos(float v) { return (v != v) ? v : __fadd_rn(fmaxf(v, 0.0f), 0.0f); }

struct Smem {
    float w1[kC1 * 4];                  // [32][4] folded layer-1 weights
    float b1[kC1];
    float b2[kC2];
    unsigned wmax[4][kC2];              // per-warp column maxima (bit patterns)
    unsigned long long bar;             // mbarrier: the centre's MMAs have completed
    uint32_t tmem;
};

__global__ void __launch_bounds__(kThreads)
local_ppf_mlp_max_kernel(const float* __restrict__ pc, const float* __restrict__ pn, const float* __restrict__ cc,
                         const float* __restrict__ cn, const int* __restrict__ nbr, int N, int M, long long centres,
                         const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                         const float* __restrict__ b2, float* __restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char lm_smem[];
    unsigned char* a_hi = lm_smem;
    unsigned char* a_lo = a_hi + kAPlane;
    unsigned char* b_hi = a_lo + kAPlane;
    unsigned char* b_lo = b_hi + kBPlane;
    Smem* S = reinterpret_cast<Smem*>(b_lo + kBPlane);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    // ---- once per CTA: weights, barrier, tensor memory
    for (int e = t; e < kC1 * 4; e += kThreads) S->w1[e] = w1[e];
    if (t < kC1) S->b1[t] = b1[t];
    if (t < kC2) S->b2[t] = b2[t];
    for (int e = t; e < kC2 * kC1; e += kThreads) {              // W2 [64][32] -> hi / lo planes in operand layout
        const int row = e / kC1, k = e - row * kC1;
        const float v = w2[e];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const uint32_t o = op_offset(kC2, row, k);
        *reinterpret_cast<float*>(b_hi + o) = h;
        *reinterpret_cast<float*>(b_lo + o) = __fsub_rn(v, h);
    }
    if (t == 0) {
        mbar_init(ri_smem_u32(&S->bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(ri_smem_u32(&S->tmem)), "r"((uint32_t)kC2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ri_fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem;
    const uint32_t a_hi_s = ri_smem_u32(a_hi), a_lo_s = ri_smem_u32(a_lo), b_hi_s = ri_smem_u32(b_hi), b_lo_s = ri_smem_u32(b_lo);
    const uint32_t row_off = (uint32_t)((t >> 3) * kSBO + (t & 7) * 16);      // this thread's row inside a 16-k chunk

    uint32_t phase = 0;
    for (long long g = blockIdx.x; g < centres; g += gridDim.x) {
        const int b = (int)(g / M), m = (int)(g - (long long)b * M);
        // ---- point-pair features of (centre m, neighbour t): local_ppf_kernel's arithmetic
        const float* PC = pc + (size_t)b * 3 * N;
        const float* PN = pn + (size_t)b * 3 * N;
        const float* CC = cc + (size_t)b * 3 * M;
        const float* CN = cn + (size_t)b * 3 * M;
        int j = nbr[((size_t)b * M + m) * kU + t];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        const float c0 = CC[m], c1 = CC[m + M], c2 = CC[m + 2 * (size_t)M];
        const float n0 = CN[m], n1 = CN[m + M], n2 = CN[m + 2 * (size_t)M];
        const float q0 = __ldg(PC + j), q1 = __ldg(PC + j + N), q2 = __ldg(PC + j + 2 * (size_t)N);
        const float r0 = __ldg(PN + j), r1 = __ldg(PN + j + N), r2 = __ldg(PN + j + 2 * (size_t)N);
        const float d0 = __fsub_rn(c0, __fsub_rn(q0, c0)), d1 = __fsub_rn(c1, __fsub_rn(q1, c1)), d2 = __fsub_rn(c2, __fsub_rn(q2, c2));
        const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
        const float u0 = __fdiv_rn(d0, dn), u1 = __fdiv_rn(d1, dn), u2 = __fdiv_rn(d2, dn);
        const float x0 = lp_acos_clamped(lp_dot(r0, r1, r2, u0, u1, u2));
        const float x1 = lp_acos_clamped(lp_dot(n0, n1, n2, u0, u1, u2));
        const float x2 = lp_acos_clamped(lp_dot(r0, r1, r2, n0, n1, n2));
        const float x3 = dn;
        // ---- layer 1 on the CUDA cores, split, written as this thread's row of the A operand (16-byte k-cores)
#pragma unroll
        for (int kq = 0; kq < kC1 / 4; ++kq) {
            float4 h, l;
            float hv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = kq * 4 + e;
                const float4 w = *reinterpret_cast<const float4*>(&S->w1[c * 4]);
                float acc = __fmaf_rn(w.x, x0, S->b1[c]);
                acc = __fmaf_rn(w.y, x1, acc);
                acc = __fmaf_rn(w.z, x2, acc);
                acc = __fmaf_rn(w.w, x3, acc);
                hv[e] = relu_pos(acc);
            }
            h.x = __uint_as_float(__float_as_uint(hv[0]) & 0xffffe000u); l.x = __fsub_rn(hv[0], h.x);
            h.y = __uint_as_float(__float_as_uint(hv[1]) & 0xffffe000u); l.y = __fsub_rn(hv[1], h.y);
            h.z = __uint_as_float(__float_as_uint(hv[2]) & 0xffffe000u); l.z = __fsub_rn(hv[2], h.z);
            h.w = __uint_as_float(__float_as_uint(hv[3]) & 0xffffe000u); l.w = __fsub_rn(hv[3], h.w);
            const uint32_t o = (uint32_t)((kq >> 2) * kU * 64) + row_off + (uint32_t)(kq & 3) * kLBO;
            *reinterpret_cast<float4*>(a_hi + o) = h;
            *reinterpret_cast<float4*>(a_lo + o) = l;
        }
        ri_fence_proxy_async_smem();                             // generic-proxy stores -> visible to the tensor core
        tc_fence_before();                                       // (and the previous centre's tcgen05.ld are done)
        __syncthreads();
        // ---- layer 2 on the tensor cores: 2 chunks x 2 K-steps x 3 split products
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ao = (uint32_t)(kc * kU * 64) + ks * 2 * kLBO, bo = (uint32_t)(kc * kC2 * 64) + ks * 2 * kLBO;
                    const uint64_t ah = smem_desc(a_hi_s + ao), al = smem_desc(a_lo_s + ao);
                    const uint64_t bh = smem_desc(b_hi_s + bo), bl = smem_desc(b_lo_s + bo);
                    tc_mma_tf32(tmem, al, bh, kIdesc, (kc | ks) != 0);          // small terms first
                    tc_mma_tf32(tmem, ah, bl, kIdesc, 1);
                    tc_mma_tf32(tmem, ah, bh, kIdesc, 1);
                }
            }
            tc_commit(ri_smem_u32(&S->bar));
        }
        mbar_wait(ri_smem_u32(&S->bar), phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue: this thread's row (TMEM lane t), bias + ReLU, maximum over the 128 rows
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32), v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float y = relu_pos(__fadd_rn(__uint_as_float(v[c]), S->b2[half * 32 + c]));
                const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                if (lane == 0) S->wmax[warp][half * 32 + c] = mx;
            }
        }
        __syncthreads();
        if (t < kC2) {
            const unsigned mx = max(max(S->wmax[0][t], S->wmax[1][t]), max(S->wmax[2][t], S->wmax[3][t]));
            out[((size_t)b * kC2 + t) * M + m] = __uint_as_float(mx);
        }
        // (the next centre's barrier before the MMAs also orders these reads of wmax before its rewrite)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kC2) : "memory");
}

}  // namespace

// points_coords / points_normals [B,3,N], centers_coords / centers_normals [B,3,M], neighbors [B,M,128] (ri_ball_query_f32),
// w1 [32,4], b1 [32], w2 [64,32], b2 [64] (1x1 convolutions with their BatchNorm folded in) -> out [B,64,M]
extern "C" int ri_local_ppf_mlp_max_f32(const float* points_coords, const float* points_normals, const float* centers_coords,
                                        const float* centers_normals, const int* neighbors, int B, int N, int M, int U,
                                        const float* w1, const float* b1, int C1, const float* w2, const float* b2, int C2,
                                        float* out, void* stream)
{
    if (B < 0 || N < 0 || M < 0 || U < 0 || w1 == nullptr || b1 == nullptr || w2 == nullptr || b2 == nullptr) return RI_ERR_BAD_ARG;
    if (U != kU || C1 != kC1 || C2 != kC2 || B > 65535) return RI_ERR_UNSUPPORTED;      // the shipped models' shape
    if (B == 0 || M == 0) return RI_OK;
    if (N == 0) return RI_ERR_BAD_ARG;
    const size_t smem = 2 * (size_t)kAPlane + 2 * (size_t)kBPlane + sizeof(Smem) + 1024;
    RI_KERNEL_SETUP(local_ppf_mlp_max_kernel, true, -1);
    const long long centres = (long long)B * M;
    long long grid = 4LL * ri_num_sms();
    if (grid > centres) grid = centres;
    local_ppf_mlp_max_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(
        points_coords, points_normals, centers_coords, centers_normals, neighbors, N, M, centres, w1, b1, w2, b2, out);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
