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
//     three shuffles away, its 32 x NB distances are 6 two-wide FP instructions per PAIR of references (FADD2 / FMUL2 /
//     FFMA2) with no memory access at all, and are parked in the lane's own shared-memory words;
//   * the k-th smallest of the 32 per-lane minima is an upper bound T of the true k-th distance — k different candidates
//     are <= T — and a tight one: on average 1.5 k candidates pass (k = 20: 30, 99th percentile 42).  It is found by
//     COUNTING (every lane reads the 32 minima from shared memory and counts those at or below its own) — a bitonic sort
//     across the lanes is fewer instructions but a chain of 15 dependent shuffles, and with 3 warps per scheduler the
//     kernel is latency-bound (measured: 40.7 us with the two sorts, issue slots 47 % used);
//   * the candidates with d <= T are compacted into shared memory (one shared-memory atomic per lane: their order does not
//     matter) as 64-bit keys (distance bits << 32 | reference index); the rank of a key among them — counted the same way —
//     is its output slot, written straight to global memory.  The key order is the reference's (distance, lower index
//     first) whatever the arrival order.  A query with more survivors than the key buffer holds streams its distance
//     rows through a bitonic sort + merge instead (same result).
// Work is cut into equal runs of consecutive (cloud, query) pairs, one run per warp of a grid that fills the machine
// once (3 CTAs of 4 warps per SM), so every SM finishes at the same time; a run that crosses a cloud boundary reloads
// its reference registers.
#include "ri_common.cuh"

namespace {

constexpr float kUndefDist = 10000.0f;                 // knn/knn.cuh:3 (UNDEFINE_VALUE)
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kNoKey = ~0ull;
constexpr int kWarpsPerCta = 4;

// ascending bitonic sort of one key per lane
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
// (q - r)^2 summed in the reference's order for two references at once: fma(dz,dz,fma(dy,dy,dx*dx)).  The references are
// held NEGATED (nr = -r): q + (-r) is q - r bit for bit, and FADD2 has no negate modifier (sub.f32x2 costs two extra FADDs).
__device__ __forceinline__ unsigned long long sqdist3x2(unsigned long long qx, unsigned long long qy, unsigned long long qz,
                                                        unsigned long long nrx, unsigned long long nry, unsigned long long nrz)
{
    unsigned long long dx, dy, dz, d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(qx), "l"(nrx));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(qy), "l"(nry));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(qz), "l"(nrz));
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
// n += (a <= b) / n += (a < b) as compare + predicated add (two / three SASS instructions, no select)
__device__ __forceinline__ void count_le(int& n, float a, float b)
{
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(n) : "f"(a), "f"(b));
}
__device__ __forceinline__ void count_lt(int& n, unsigned long long a, unsigned long long b)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(n) : "l"(a), "l"(b));
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

constexpr int kKeyCap = 128;      // compacted survivors per query kept in shared memory (typical: 30; beyond: streamed)

// shared memory of one warp, in 8-byte words: keys [kKeyCap] | distances [16 * NB] | lane minima (32 floats) | counter
template <int NB>
struct WarpSmem {
    static constexpr int kKeys = 0, kDist = kKeyCap, kMin = kKeyCap + 16 * NB, kCount = kMin + 16, kWords = kCount + 2;
};

// queries [B,3,n], refs [B,3,m] -> dist / idx [B,k,n].  NB = references per lane (m <= 32 * NB), even.
template <int NB>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
knn3_warp_kernel(const float* __restrict__ queries, const float* __restrict__ refs, int n, int m, int k, long long total,
                 int qpw, float* __restrict__ dist, int* __restrict__ idx)
{
    extern __shared__ __align__(16) unsigned long long smem[];
    using W = WarpSmem<NB>;
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler too
    unsigned long long* const wbase = smem + (size_t)wid * W::kWords;
    unsigned long long* const keys = wbase + W::kKeys;
    unsigned long long* const d2 = wbase + W::kDist;                      // [NB/2][32] pairs (d[2p], d[2p+1]) of lane
    const float* const df = reinterpret_cast<const float*>(d2);
    float* const smin = reinterpret_cast<float*>(wbase + W::kMin);
    int* const scount = reinterpret_cast<int*>(wbase + W::kCount);
    const long long g0 = ((long long)blockIdx.x * kWarpsPerCta + wid) * qpw;
    const long long g1 = g0 + qpw < total ? g0 + qpw : total;
    if (g0 >= g1) return;
    const float kBelowUndef = __uint_as_float(__float_as_uint(kUndefDist) - 1u);   // largest float < 10000
    if (lane == 0) *scount = 0;
    for (int e = lane; e < kKeyCap; e += 32) keys[e] = kNoKey;            // invariant: slots beyond the survivors hold kNoKey

    unsigned long long RX[NB / 2], RY[NB / 2], RZ[NB / 2];                // the cloud's references, NEGATED, two per register pair
    int cur_b = -1;
    int b = (int)(g0 / n), qi = (int)(g0 % n);       // the run's current (cloud, query): warp-uniform, advanced by hand
    float cqx = 0.f, cqy = 0.f, cqz = 0.f;
    for (long long g = g0; g < g1; ++g) {
        const int j = (int)(g - g0) & 31;
        if (j == 0) {                                // next 32 queries of the run: lane l fetches query g + l (coalesced)
            const long long t = (long long)qi + lane;
            if (g + lane < g1) {
                const float* Q = queries + ((size_t)b + (size_t)(t / n)) * 3 * n + (int)(t % n);
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
                RX[p] = pack2(in0 ? -R[t0] : INFINITY, in1 ? -R[t1] : INFINITY);
                RY[p] = pack2(in0 ? -R[t0 + m] : INFINITY, in1 ? -R[t1 + m] : INFINITY);
                RZ[p] = pack2(in0 ? -R[t0 + 2 * (size_t)m] : INFINITY, in1 ? -R[t1 + 2 * (size_t)m] : INFINITY);
            }
            cur_b = b;
        }
        // ---- all distances of this query, parked in the lane's own shared-memory words; the lane's minimum
        const unsigned long long Q0 = pack2(qx, qx), Q1 = pack2(qy, qy), Q2 = pack2(qz, qz);
        float lmin = INFINITY;
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) {
            const unsigned long long dd = sqdist3x2(Q0, Q1, Q2, RX[p], RY[p], RZ[p]);
            float da, db;
            unpack2(dd, da, db);
            lmin = fminf(lmin, fminf(da, db));       // fminf drops a NaN distance
            d2[p * 32 + lane] = dd;
        }
        smin[lane] = lmin;
        __syncwarp();
        // ---- T = k-th smallest of the 32 lane minima: k distinct candidates are <= T, so the true k-th distance is too.
        //      By counting, not sorting (no chain of dependent shuffles): a lane whose minimum has >= k minima at or below it
        //      holds a value >= the k-th smallest, the smallest such value IS the k-th smallest.
        int le0 = 0, le1 = 0, le2 = 0, le3 = 0;      // four accumulators: no serial chain of adds
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float4 v = reinterpret_cast<const float4*>(smin)[t];
            count_le(le0, v.x, lmin); count_le(le1, v.y, lmin); count_le(le2, v.z, lmin); count_le(le3, v.w, lmin);
        }
        const int le = (le0 + le1) + (le2 + le3);
        float T = __uint_as_float(__reduce_min_sync(kFull, le >= k ? __float_as_uint(lmin) : 0x7f800000u));
        T = fminf(T, kBelowUndef);                   // the reference admits d < 10000 only (this also cuts the +inf padding)
        // ---- the candidates with d <= T: a bit per reference of this lane ...
        unsigned mask = 0u;
        {
            float d[NB];
#pragma unroll
            for (int p = 0; p < NB / 2; ++p) unpack2(d2[p * 32 + lane], d[2 * p], d[2 * p + 1]);
            MarkAll<0, NB>::run(mask, d, T);
        }
        // ---- ... compacted into shared memory as (distance, index) keys; their order in the buffer does not matter
        const int c = __popc(mask);
        int off = c ? atomicAdd(scount, c) : 0;
        for (unsigned mk = off + c <= kKeyCap ? mask : 0u; mk; mk &= mk - 1) {
            const int i = __ffs(mk) - 1;
            const float dv = df[(i >> 1) * 64 + 2 * lane + (i & 1)];
            keys[off++] = ((unsigned long long)__float_as_uint(dv) << 32) | (unsigned)(32 * i + lane);
        }
        __syncwarp();
        const int S = *scount;
        float* od = dist + (size_t)b * k * n + qi;
        int* oi = idx + (size_t)b * k * n + qi;
        if (S <= kKeyCap) {
            // ---- rank of every key among the survivors = its place in the output (keys are distinct: the index is part of
            //      them), again by counting: every lane reads all keys (broadcast loads), no dependent shuffles
            for (int e0 = 0; e0 < S; e0 += 32) {                 // warp-uniform trip count: one pass per 32 survivors
                const int e = e0 + lane;
                const unsigned long long x = keys[e];             // slots >= S hold kNoKey
                int r0 = 0, r1 = 0;
                for (int t = 0; t < S; t += 4) {                  // the buffer is padded with kNoKey: no tail
                    const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(keys + t);
                    const ulonglong2 z = *reinterpret_cast<const ulonglong2*>(keys + t + 2);
                    count_lt(r0, y.x, x); count_lt(r1, y.y, x); count_lt(r0, z.x, x); count_lt(r1, z.y, x);
                }
                const int rank = r0 + r1;
                if (e < S && rank < k) {
                    od[(size_t)rank * n] = __uint_as_float((unsigned)(x >> 32));
                    oi[(size_t)rank * n] = (int)(unsigned)(x & 0xffffffffu);
                }
            }
            if (lane >= S && lane < k) { od[(size_t)lane * n] = kUndefDist; oi[(size_t)lane * n] = 0; }
        } else {
            // more survivors than the key buffer holds (one lane's share of the references holds most of the near ones, e.g.
            // an index order with period 32): stream the distance rows, 32 candidates (one per lane) at a time, through a
            // bitonic sort + merge that keeps the 32 smallest keys in the lanes
            unsigned long long L = kNoKey;
            for (int i = 0; i < NB; ++i) {
                const float dv = df[(i >> 1) * 64 + 2 * lane + (i & 1)];
                const unsigned long long kth = __shfl_sync(kFull, L, k - 1);
                unsigned long long x = dv <= T ? (((unsigned long long)__float_as_uint(dv) << 32) | (unsigned)(32 * i + lane)) : kNoKey;
                if (!__any_sync(kFull, x < kth)) continue;
                x = warp_sort_u64(x, lane);
                const unsigned long long rev = __shfl_sync(kFull, x, 31 - lane);
                L = warp_merge_u64(rev < L ? rev : L, lane);
            }
            if (lane < k) {
                od[(size_t)lane * n] = L == kNoKey ? kUndefDist : __uint_as_float((unsigned)(L >> 32));
                oi[(size_t)lane * n] = L == kNoKey ? 0 : (int)(unsigned)(L & 0xffffffffu);
            }
        }
        __syncwarp();                                // every lane has read the counter and the keys: reuse them
        if (lane == 0) *scount = 0;
        for (int e = lane; e < S && e < kKeyCap; e += 32) keys[e] = kNoKey;
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
    const size_t smem = (size_t)kWarpsPerCta * WarpSmem<NB>::kWords * sizeof(unsigned long long);
    RI_KERNEL_SETUP(knn3_warp_kernel<NB>, false, ri_step_carveout_percent());
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
