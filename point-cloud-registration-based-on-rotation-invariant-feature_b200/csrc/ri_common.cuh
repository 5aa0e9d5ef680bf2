// ri_common.cuh — shared device helpers for the sm_100a kernels behind include/ri_b200.h.
//
// Arithmetic pins.  Where the contract is bit-exactness against the reference kernels (voxel indices,
// counts, KNN order, devox corner indices) the fp32/fp64 operation ORDER is spelled with explicit
// round-to-nearest intrinsics (__fmaf_rn, __fmul_rn, __fdiv_rn, __fsqrt_rn, __dadd_rn ...), so it cannot
// drift with compiler flags or contraction heuristics.  The order itself was read from the sm_100a PTX of
// the reference sources (/root/reference/PVCNN/modules/functional/src/**); each helper cites the lines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>

#define RI_OK 0
#define RI_ERR_BAD_ARG (-1)
#define RI_ERR_WORKSPACE (-2)
#define RI_ERR_UNSUPPORTED (-3)

#define RI_SM_COUNT_FALLBACK 148

// acos(-1.0) as the reference evaluates it on the device (`#define PI acos(-1.0)`,
// spherical_voxelization/spherical_vox.cu:5): the inlined f64 acos returns 0x400921FB54442D18 == M_PI.
#define RI_PI 3.14159265358979311600e+00

#define RI_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

// ---- process-wide state of the library: everything below is read-only after its first use and safe to reach from several
// host threads and several devices of one process (the reference trains with single-process nn.DataParallel) -----------

// Environment knobs (experiments and tests — none changes a result).  Read ONCE per process: the
// initialisation of a function-local static is thread-safe, the launch paths never call getenv().
struct RiEnv {
    int carveout_pct;                       // RI_CARVEOUT_PCT       shared-memory carveout preference of the step's kernels
    int devox_stream;                       // RI_DEVOX_STREAM       -1 unset, 0 gather form, 1 streaming form wherever it can run
    int devox_tile_kb, devox_ring_kb, devox_pad_kb, devox_dbg_skip;
    int fill_ring, fill_ctas, fill_pad_kb;  // RI_FILL_*             grid writer: slots per warp, CTAs per SM, shared-memory padding
    int fill_group;                         // RI_FILL_GROUP         grid writer: planes per work item
    int fill_warps, fill_listcap, fill_tile_cells;   // RI_FILL_WARPS / _LISTCAP / _TILE   writer warps per CTA, staged cells per tile, cells per tile
    int vox_atomic;                         // RI_VOX_ATOMIC         scan-sized clouds: memset + float atomics
    int ppf_maxl1;                          // RI_PPF_MAXL1          packed PPF kernel keeps the default (max-L1) split
    int match_pair;                         // RI_MATCH_PAIR         -1 unset, 0 single-CTA form, 1 CTA-pair form
    int match_dbg;                          // RI_MATCH_DBG          GEMM only, %globaltimer stamps in the workspace
    int match_tma;                          // RI_MATCH_TMA          0: never take the tensor-map (no-image) path of the indices-only matcher
};
// One instance per process, defined in abi.cu; initialised on first use (thread-safe), read-only on every launch path.
// ri_debug_set_knob() (tests, tools) overwrites a field — not while launches are in flight on other threads.
const RiEnv& ri_env();

// SM count of the CURRENT device, cached per device ordinal.
static inline int ri_num_sms()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RI_SM_COUNT_FALLBACK;
    const int slot = dev & 63;
    int n = cached[slot].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = RI_SM_COUNT_FALLBACK;
        cached[slot].store(n, std::memory_order_relaxed);
    }
    return n;
}

// L1 / shared-memory split of the kernels that are meant to run NEXT TO the grid writer (vox_fill: a 96 KB ring per SM).
// An SM cannot host CTAs of kernels configured for different carveouts at the same time — it drains first — which
// silently serialises branches that were meant to overlap (measured on a B200: k-NN next to the grid writer 154 us with
// mismatched carveouts, 108 us matched; k-NN next to the devoxelizer 285 us against 78 + 54 us one after the other).
// So the k-NN / PPF branch and the prefix of the voxel branch all ask for the writer's max-shared split.  The gather
// devoxelizer must NOT: its gathers use L1 as their miss buffer and its speed follows the L1 size (54 us with the
// default max-L1 split, 73 / 88 / 118 / 231 us with 100 / 132 / 164 / 228 KB of shared memory), so it keeps the default
// and is scheduled after the max-shared kernels have left the SMs.
static inline int ri_step_carveout_percent() { return ri_env().carveout_pct; }

// Function attributes are PER DEVICE (and a kernel's dynamic shared memory above 48 KB is an opt-in): they are set the
// first time a kernel is launched on a device, not once per process and not on every launch.  One lock-free table per
// translation unit, keyed by the kernel's address: slot = {kernel, bit mask of the devices done}.  Two threads racing on
// the same (kernel, device) both set the same values — harmless.
//   smem_optin  raise cudaFuncAttributeMaxDynamicSharedMemorySize to everything the device allows (minus the kernel's
//               static shared memory), so any later launch size is legal
//   carveout    >= 0: cudaFuncAttributePreferredSharedMemoryCarveout
static inline cudaError_t ri_kernel_setup(const void* kernel, bool smem_optin, int carveout)
{
    constexpr int kSlots = 128;
    static std::atomic<const void*> key[kSlots];
    static std::atomic<unsigned long long> done[kSlots];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    int slot = (int)(((uintptr_t)kernel >> 4) % kSlots);
    for (int probe = 0; probe < kSlots; ++probe, slot = (slot + 1) % kSlots) {
        const void* cur = key[slot].load(std::memory_order_acquire);
        if (cur == kernel) break;
        if (cur == nullptr) {
            const void* expected = nullptr;
            if (key[slot].compare_exchange_strong(expected, kernel, std::memory_order_acq_rel) || expected == kernel) break;
        }
    }
    const bool tracked = key[slot].load(std::memory_order_acquire) == kernel;       // table full: set on every launch
    if (tracked && (done[slot].load(std::memory_order_acquire) & bit)) return cudaSuccess;
    if (smem_optin) {
        int optin = 0;
        cudaFuncAttributes fa;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return e;
    }
    if (carveout >= 0) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        if (e != cudaSuccess) return e;
    }
    if (tracked) done[slot].fetch_or(bit, std::memory_order_release);
    return cudaSuccess;
}
#define RI_KERNEL_SETUP(kernel, smem_optin, carveout)                                              \
    do {                                                                                           \
        cudaError_t e__ = ri_kernel_setup((const void*)(kernel), (smem_optin), (carveout));        \
        if (e__ != cudaSuccess) return (int)e__;                                                   \
    } while (0)
template <typename K>
static inline void ri_prefer_step_carveout(K kernel)
{
    ri_kernel_setup((const void*)kernel, false, ri_step_carveout_percent());
}

// x*x' + y*y' + z*z' as nvcc contracts it for the reference: fma(z,z', fma(x,x', y*y')).
// (PTX of ppf.cu:56-83 and spherical_vox.cu:37: mul.f32 on the y term, then two fma.rn.f32.)
__device__ __forceinline__ float ri_dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}

// Spherical coordinates of a normalised point, exactly as spherical_vox.cu:34-56 and
// spherical_trilinear_devox.cu:48-65 evaluate them (mixed f32/f64, libdevice acosf/atanf).
// Returns false for an "undefined" point (ind = -1).
__device__ __forceinline__ bool ri_sph_coords(float x, float y, float z, int r,
                                              float& gama, float& alpha, float& beta)
{
    gama = __fsqrt_rn(ri_dot3(x, y, z, x, y, z));
    if (gama == 0.0f || gama >= 1.0f) return false;
    const float t = __fdiv_rn(z, gama);
    if (t > 1.0f || t < -1.0f) return false;
    beta = acosf(t);
    if ((double)beta >= RI_PI) return false;
    if (x == 0.0f && y != 0.0f) {
        alpha = __double2float_rn(__dmul_rn(__dmul_rn((double)__fdiv_rn(y, fabsf(y)), RI_PI), 0.5));
    } else if (x == 0.0f && y == 0.0f) {
        alpha = 0.0f;
    } else {
        const float sgn = __fsub_rn(1.0f, __fdiv_rn(x, fabsf(x)));
        alpha = __double2float_rn(__fma_rn(__dmul_rn(RI_PI, (double)sgn), 0.5, (double)atanf(__fdiv_rn(y, x))));
    }
    alpha = __double2float_rn(__dadd_rn(__ddiv_rn(RI_PI, (double)r), (double)alpha));
    if (alpha < 0.0f) alpha = __double2float_rn(__fma_rn(RI_PI, 2.0, (double)alpha));
    return true;
}

// Spherical cell of a point: spherical_vox.cu:59-65.  -1 if undefined.
__device__ __forceinline__ int ri_sph_cell(float x, float y, float z, int r)
{
    float g, a, be;
    if (!ri_sph_coords(x, y, z, r, g, a, be)) return -1;
    const float rf = (float)r;
    int gx = (int)floorf(__fmul_rn(g, rf));
    int gy = (int)floor(__ddiv_rn((double)__fmul_rn(__fmul_rn(a, rf), 0.5f), RI_PI));
    int gz = (int)floor(__ddiv_rn((double)__fmul_rn(be, rf), RI_PI));
    gx = gx >= r ? r - 1 : gx;
    gy = gy >= r ? r - 1 : gy;
    gz = gz >= r ? r - 1 : gz;
    return gx * r * r + gy * r + gz;
}

// Corner weights / indices shared by both devoxelizers (interpolate/trilinear_devox.cu:46-76).
__device__ __forceinline__ void ri_corners(float d1a, float d1b, float d1c, int lo_a, int lo_b, int lo_c,
                                           int r, int r2, int (&id)[8], float (&w)[8])
{
    const float d0a = __fsub_rn(1.0f, d1a), d0b = __fsub_rn(1.0f, d1b), d0c = __fsub_rn(1.0f, d1c);
    const float w00 = __fmul_rn(d0a, d0b), w01 = __fmul_rn(d0a, d1b);
    const float w10 = __fmul_rn(d1a, d0b), w11 = __fmul_rn(d1a, d1b);
    w[0] = __fmul_rn(w00, d0c); w[1] = __fmul_rn(w00, d1c);
    w[2] = __fmul_rn(w01, d0c); w[3] = __fmul_rn(w01, d1c);
    w[4] = __fmul_rn(w10, d0c); w[5] = __fmul_rn(w10, d1c);
    w[6] = __fmul_rn(w11, d0c); w[7] = __fmul_rn(w11, d1c);
    const int ha = d1a > 0.0f ? r2 : 0, hb = d1b > 0.0f ? r : 0, hc = d1c > 0.0f ? 1 : 0;
    id[0] = lo_a * r2 + lo_b * r + lo_c;
    id[1] = id[0] + hc;
    id[2] = id[0] + hb;
    id[3] = id[2] + hc;
    id[4] = id[0] + ha;
    id[5] = id[4] + hc;
    id[6] = id[4] + hb;
    id[7] = id[6] + hc;
}

// ---- PTX wrappers: bulk async copy (TMA engine, SASS UBLKCP) and proxy fence -------------------------
__device__ __forceinline__ uint32_t ri_smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ri_fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void ri_bulk_store(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(ri_smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ri_bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void ri_bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void ri_bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
