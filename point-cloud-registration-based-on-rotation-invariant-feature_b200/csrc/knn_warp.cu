// knn_warp.cu — exact brute-force k-NN for small clouds (c == 3, m <= 1024 references, k <= 32): one WARP per query,
// selection by an exact threshold instead of a running sorted list.  sm_100a.
//
// Same contract as knn.cu (KnnKernel, /root/reference/PVCNN/modules/functional/src/knn/knn.cu:5-49): squared L2
// d = fma(dz,dz,fma(dy,dy,dx*dx)) with d_p = query_p - ref_p; a candidate needs d < 10000; result = the first k of a
// stable sort by distance (lower reference index first), unfilled slots (10000.0f, 0).
//
// Why not one thread per query (knn.cu, 78 us for 32 x 1024 x 1024, k = 20): its cost is the running top-k list — a
// sorted insertion is ~130 predicated instructions for the whole warp whenever ONE lane has a survivor, ~250 times per
// warp — and 32 768 queries are only 1 024 warps, 1.7 per scheduler.  Here:
//   * lane l keeps the references l, 32+l, 64+l, ... of the cloud in REGISTERS (3 x NB floats, NB = m/32 <= 32); a query is
//     three shuffles away, its 32 x NB distances are 6 FP instructions each with no memory access at all, and stay in
//     registers too;
//   * the k-th smallest of the 32 per-lane minima (one 15-step bitonic sort of floats across the warp) is an upper bound
//     T of the true k-th distance — k different candidates are <= T — and a tight one: on average ~1.5 k candidates pass;
//   * the candidates with d <= T are compacted into shared memory (per-lane counts, one warp scan) as 64-bit keys
//     (distance bits << 32 | reference index), sorted 32 at a time by a bitonic network across the lanes and merged; lane s
//     writes neighbour s.  The key order is the reference's (distance, lower index first) whatever the arrival order.
// Work is cut into equal runs of consecutive (cloud, query) pairs, one run per warp of a grid that fills the machine
// once (3 CTAs of 4 warps per SM), so every SM finishes at the same time; a run that crosses a cloud boundary reloads
// its reference registers.
#include "ri_common.cuh"

namespace {

constexpr float kUndefDist = 10000.0f;                 // knn/knn.cuh:3 (UNDEFINE_VALUE)
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kNoKey = ~0ull;
constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, float rx, float ry, float rz)
{
    const float dx = __fsub_rn(qx, rx), dy = __fsub_rn(qy, ry), dz = __fsub_rn(qz, rz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// ascending bitonic sort of one value per lane
__device__ __forceinline__ float warp_sort_f32(float v, int lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(kFull, v, j);
            const bool take_min = ((lane & kk) == 0) == ((lane & j) == 0);
            v = take_min ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_sort_u64(unsigned long long v, int lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, v, j);
            const bool take_min = ((lane & kk) == 0) == ((lane & j) == 0);
            v = ((o < v) == take_min) ? o : v;
        }
    }
    return v;
}
// v: a bitonic sequence across the lanes -> ascending
__device__ __forceinline__ unsigned long long warp_merge_u64(unsigned long long v, int lane)
{
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFull, v, j);
        const bool take_min = (lane & j) == 0;
        v = ((o < v) == take_min) ? o : v;
    }
    return v;
}

// packed pairs of floats: sm_100a has two-wide fp32 add / mul / fma (SASS FADD2 / FMUL2 / FFMA2), IEEE per component
__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// (q - r)^2 summed in the reference's order for two references at once: fma(dz,dz,fma(dy,dy,dx*dx))
__device__ __forceinline__ unsigned long long sqdist3x2(unsigned long long qx, unsigned long long qy, unsigned long long qz,
                                                        unsigned long long rx, unsigned long long ry, unsigned long long rz)
{
    unsigned long long dx, dy, dz, d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(qx), "l"(rx));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(qy), "l"(ry));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(qz), "l"(rz));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dx));
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dy), "l"(d));
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dz), "l"(d));
    return d;
}
template <unsigned BIT>
__device__ __forceinline__ void mark_le(unsigned& mask, float d, float T)
{
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(mask) : "f"(d), "f"(T), "n"(1u << BIT));
}
template <int I, int N>
struct MarkAll {
    static __device__ __forceinline__ void run(unsigned& mask, const float (&d)[N], float T)
    {
        mark_le<I>(mask, d[I], T);
        MarkAll<I + 1, N>::run(mask, d, T);
    }
};
template <int N>
struct MarkAll<N, N> {
    static __device__ __forceinline__ void run(unsigned&, const float (&)[N], float) {}
};

// queries [B,3,n], refs [B,3,m] -> dist / idx [B,k,n].  NB = references per lane (m <= 32 * NB), even.
template <int NB>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
knn3_warp_kernel(const float* __restrict__ queries, const float* __restrict__ refs, int n, int m, int k, long long total,
                 int qpw, float* __restrict__ dist, int* __restrict__ idx)
{
    extern __shared__ unsigned long long smem[];     // per warp: 32 * NB keys, then 32 * NB distances
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler too
    unsigned long long* mykeys = smem + (size_t)wid * (32 * NB + 16 * NB);
    unsigned long long* myd2 = mykeys + 32 * NB;                          // [NB/2][32] pairs (d[2p], d[2p+1]) of lane
    const float* mydf = reinterpret_cast<const float*>(myd2);
    const long long g0 = ((long long)blockIdx.x * kWarpsPerCta + wid) * qpw;
    const long long g1 = g0 + qpw < total ? g0 + qpw : total;
    if (g0 >= g1) return;
    const float kBelowUndef = __uint_as_float(__float_as_uint(kUndefDist) - 1u);   // largest float < 10000

    unsigned long long RX[NB / 2], RY[NB / 2], RZ[NB / 2];
    int cur_b = -1;
    int b = (int)(g0 / n), qi = (int)(g0 % n);       // the run's current (cloud, query): warp-uniform, advanced by hand
    float cqx = 0.f, cqy = 0.f, cqz = 0.f;
    for (long long g = g0; g < g1; ++g) {
        const int j = (int)(g - g0) & 31;
        if (j == 0) {                                // next 32 queries of the run: lane l fetches query g + l
            const long long t = (long long)qi + lane;
            const long long bj = b + t / n;
            const int ij = (int)(t % n);
            if (g + lane < g1) {
                const float* Q = queries + (size_t)bj * 3 * n + ij;
                cqx = Q[0]; cqy = Q[n]; cqz = Q[2 * (size_t)n];
            }
        }
        const float qx = __shfl_sync(kFull, cqx, j), qy = __shfl_sync(kFull, cqy, j), qz = __shfl_sync(kFull, cqz, j);
        if (b != cur_b) {                            // the run entered another cloud: reload the reference registers
            const float* R = refs + (size_t)b * 3 * m;
#pragma unroll
            for (int p = 0; p < NB / 2; ++p) {
                const int t0 = 64 * p + lane, t1 = t0 + 32;
                const bool in0 = t0 < m, in1 = t1 < m;
                RX[p] = pack2(in0 ? R[t0] : INFINITY, in1 ? R[t1] : INFINITY);
                RY[p] = pack2(in0 ? R[t0 + m] : INFINITY, in1 ? R[t1 + m] : INFINITY);
                RZ[p] = pack2(in0 ? R[t0 + 2 * (size_t)m] : INFINITY, in1 ? R[t1 + 2 * (size_t)m] : INFINITY);
            }
            cur_b = b;
        }
        // ---- all distances of this query (kept in registers, copied to shared memory for the indexed re-read below)
        const unsigned long long Q0 = pack2(qx, qx), Q1 = pack2(qy, qy), Q2 = pack2(qz, qz);
        float d[NB];
        float lmin = INFINITY;
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) {
            const unsigned long long dd = sqdist3x2(Q0, Q1, Q2, RX[p], RY[p], RZ[p]);
            unpack2(dd, d[2 * p], d[2 * p + 1]);
            lmin = fminf(lmin, fminf(d[2 * p], d[2 * p + 1]));     // fminf drops a NaN distance
            myd2[p * 32 + lane] = dd;
        }
        // ---- T = k-th smallest lane minimum: k distinct candidates are <= T, so the true k-th distance is too
        float T = __shfl_sync(kFull, warp_sort_f32(lmin, lane), k - 1);
        T = fminf(T, kBelowUndef);                   // the reference admits d < 10000 only (also cuts the +inf padding)
        // ---- the candidates with d <= T: a bit per reference of this lane, then compacted into shared memory as keys
        unsigned mask = 0u;
        MarkAll<0, NB>::run(mask, d, T);
        const int c = __popc(mask);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        const int S = __shfl_sync(kFull, inc, 31);
        int off = inc - c;
        while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1;
            const float dv = mydf[(i >> 1) * 64 + 2 * lane + (i & 1)];
            mykeys[off++] = ((unsigned long long)__float_as_uint(dv) << 32) | (unsigned)(32 * i + lane);
        }
        __syncwarp();
        // ---- sort them, 32 at a time, keeping the 32 smallest keys in the lanes
        unsigned long long L = warp_sort_u64(lane < S ? mykeys[lane] : kNoKey, lane);
        for (int s0 = 32; s0 < S; s0 += 32) {
            const unsigned long long kth = __shfl_sync(kFull, L, k - 1);
            unsigned long long x = s0 + lane < S ? mykeys[s0 + lane] : kNoKey;
            if (!__any_sync(kFull, x < kth)) continue;
            x = warp_sort_u64(x, lane);
            const unsigned long long rev = __shfl_sync(kFull, x, 31 - lane);
            L = warp_merge_u64(rev < L ? rev : L, lane);
        }
        __syncwarp();                                // the shared lists are rewritten by the next query
        if (lane < k) {
            const size_t o = ((size_t)b * k + lane) * n + qi;
            dist[o] = L == kNoKey ? kUndefDist : __uint_as_float((unsigned)(L >> 32));
            idx[o] = L == kNoKey ? 0 : (int)(unsigned)(L & 0xffffffffu);
        }
        if (++qi == n) { qi = 0; ++b; }
    }
}

template <int NB>
int launch_nb(const float* queries, const float* refs, int B, int n, int m, int k, float* dist, int* idx, cudaStream_t st)
{
    const long long total = (long long)B * n;
    // one wave: 3 CTAs per SM; at least 8 queries per warp so that a reference reload is amortised
    long long warps = (long long)ri_num_sms() * 3 * kWarpsPerCta;
    long long qpw = (total + warps - 1) / warps;
    if (qpw < 8) qpw = 8;
    if (qpw > 0x3fffffff) return RI_ERR_UNSUPPORTED;
    warps = (total + qpw - 1) / qpw;
    const int ctas = (int)((warps + kWarpsPerCta - 1) / kWarpsPerCta);
    const size_t smem = (size_t)kWarpsPerCta * (32 * NB + 16 * NB) * sizeof(unsigned long long);
    cudaFuncSetAttribute(knn3_warp_kernel<NB>, cudaFuncAttributePreferredSharedMemoryCarveout, ri_step_carveout_percent());
    knn3_warp_kernel<NB><<<ctas, 32 * kWarpsPerCta, smem, st>>>(queries, refs, n, m, k, total, (int)qpw, dist, idx);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

}  // namespace

// Used by ri_knn_f32 (knn.cu) for c == 3, k <= 32, 1 <= m <= 1024.
int ri_launch_knn_warp(const float* queries, const float* refs, int B, int n, int m, int k, float* dist, int* idx,
                       cudaStream_t st)
{
    if (B == 0 || n == 0) return RI_OK;
    if (m <= 64) return launch_nb<2>(queries, refs, B, n, m, k, dist, idx, st);
    if (m <= 128) return launch_nb<4>(queries, refs, B, n, m, k, dist, idx, st);
    if (m <= 256) return launch_nb<8>(queries, refs, B, n, m, k, dist, idx, st);
    if (m <= 512) return launch_nb<16>(queries, refs, B, n, m, k, dist, idx, st);
    return launch_nb<32>(queries, refs, B, n, m, k, dist, idx, st);
}
