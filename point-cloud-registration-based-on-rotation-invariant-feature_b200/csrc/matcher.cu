// matcher.cu — mutual-nearest-neighbour descriptor matching on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces  find_correspondence_one_pair  (/root/reference/datasets/deepgmr_mn40.py:232-244; duplicated in
// deepgmr_partial.py:335-347, mn40_hdf.py:466-477, ...), which the registration meters call per pair on the CPU
// after a .cpu().numpy() hop (deepgmr_mn40.py:88,200):
//     diff = ||f1||^2[:,None] + ||f2||^2[None,:] - 2 f1 f2^T ;  c1 = argmin(diff, 1) ;  c2 = argmin(diff, 0)
//     mask = c2[c1] == arange(n1) ;  return arange(n1)[mask], c1[mask]
//
// B200 design.  The n1 x n2 x C contraction is the only O(n^2 C) term, so it goes on tcgen05; the n1 x n2 matrix is
// never written anywhere: the epilogue reads the fp32 accumulator tile out of TMEM, adds the norms and reduces it to
// one (distance, index) key per row and per column on the spot.
//
//   * Precision: a single tf32/bf16 pass cannot hold the 1e-5 contract (nor stable argmins), so the product is the
//     3xTF32 split  a = a_hi + a_lo  (a_hi = a with the low 13 mantissa bits cleared — representable in tf32 whatever
//     the hardware does with the low bits —, a_lo = a - a_hi, exact in fp32),  a.b ~ hi.hi + hi.lo + lo.hi, fp32
//     accumulation in TMEM.  The tensor core's accumulator adds truncate (measured 1.1e-5 of |f1|^2 + |f2|^2 on
//     well-matched 512-channel descriptors — over the contract — while the ARGMINS agree with an fp64 evaluation), so
//     this kernel decides WHO matches and match_finish reports HOW FAR, re-evaluating |f1_i - f2_j|^2 in fp32.
//   * match_prep     per cloud: squared norms (accumulated in fp64, rounded once) and a re-tiling of the raw fp32
//                    descriptors into UMMA "canonical K-major, no swizzle" core-matrix order
//                    R[kc][row/8][k16/4][row%8][4 floats]  (kc = 16-channel chunk).  An operand tile of R rows x 16
//                    channels is then ONE contiguous R*64-byte block: the GEMM kernel stages operands with plain
//                    cp.async.bulk copies (TMA engine, no tensor maps) and describes them with LBO = 128 B, SBO = 512 B.
//   * match_gemm     one CTA per 128 x 256 tile of a pair's distance matrix, 192 threads, warp-specialised:
//                      warp 4 / lane 0   TMA producer: 24 KB of raw fp32 per 16-channel stage, mbarrier complete_tx
//                      warps 0-3         converters: split the stage IN shared memory (hi written back in place, lo to
//                                        the second plane; fence.proxy.async; mbarrier arrive) — the first version
//                                        loaded pre-split planes instead, 48 KB per stage against ~810 tensor-core
//                                        cycles, i.e. 59 B/clk per SM where the L2 sustains ~42 (tensor pipe 39 %) —
//                                        and afterwards run the epilogue: tcgen05.ld 32 columns at a time,
//                                        d = (n1 + n2) - 2 acc, row argmin in registers, column argmin with redux.sync
//                                        + ballot, merged across tiles by 64-bit atomicMin on packed keys
//                                        (ordered(d) << 32 | index): lowest index wins ties, as np.argmin does
//                      warp 5            TMEM allocator; lane 0 issues tcgen05.mma.kind::tf32 (M128 N256 K8), 6 per
//                                        stage; tcgen05.commit releases the stage / publishes the accumulator
//   * match_dist     per row of every pair: unpack c1, mutual flag, and the fp32 distance of the row's match from the
//                    re-tiled descriptors;  match_compact  per pair: c2 and the ordered compaction (idx1, idx2, count).
#include "ri_common.cuh"
#include <cuda.h>
#include <string.h>

namespace {

constexpr int kTileM = 128;          // rows of f1 per CTA (TMEM lanes)
constexpr int kTileN = 256;          // rows of f2 per CTA (TMEM columns)
constexpr int kChunkK = 16;          // channels per pipeline stage (2 UMMA K-steps of 8 tf32)
constexpr int kStages = 4;
// The 'hi' operand of the 3xTF32 split is a with its low 13 mantissa bits cleared.  The tensor core ignores those bits of a
// tf32 operand (truncation — pinned by tests/test_matcher_gpu.py::test_tensor_core_truncates_tf32_operands, which fails if it
// ever rounds), so the raw fp32 tile IS the hi plane and the converters only write the lo plane: a third less shared-memory
// traffic per stage.  Set to true to write the cleared values back explicitly.
constexpr bool kWriteHi = false;
constexpr int kConvGroups = 4;                           // converter warp groups; group g owns the stages kc % kConvGroups == g.
// MUST divide kStages (static_assert below): a stage's barriers carry one phase bit, so every use of a stage has to be waited
// for by the SAME warps in order.  With 3 groups over 4 stages a group met stage 0 only every third round, its early
// try_wait(parity) was satisfied by another group's round of the same parity, it converted a tile that had not landed and
// arrived on a barrier of the wrong round: about one call in a hundred deadlocked.
constexpr int kRowBytes = kChunkK * 4;                       // 64 B of one row in one chunk
constexpr int kABytes = kTileM * kRowBytes;                  // 8 KB  (one plane)
constexpr int kBBytes = kTileN * kRowBytes;                  // 16 KB (one plane)
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;       // 48 KB
constexpr unsigned kLBO = 128, kSBO = 512;                   // see the image layout above

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct MatchWs {                     // byte offsets inside the workspace
    size_t img1, img2, nrm1, nrm2, rowkey, colkey, mutual, total;
    int n1p, n2p, Cp;
};
__host__ __device__ inline MatchWs match_ws_layout(int P, int C, int n1, int n2)
{
    MatchWs w;
    w.n1p = round_up(n1 > 0 ? n1 : 1, kTileN);        // both padded to the 256-row granule of match_prep
    w.n2p = round_up(n2 > 0 ? n2 : 1, kTileN);
    w.Cp = round_up(C > 0 ? C : 1, kChunkK);
    size_t o = 0;
    w.img1 = o; o += (size_t)P * w.n1p * w.Cp * sizeof(float);
    w.img2 = o; o += (size_t)P * w.n2p * w.Cp * sizeof(float);
    w.nrm1 = o; o += (size_t)P * w.n1p * sizeof(float);
    w.nrm2 = o; o += (size_t)P * w.n2p * sizeof(float);
    o = (o + 15) / 16 * 16;
    w.rowkey = o; o += (size_t)P * w.n1p * sizeof(unsigned long long);
    w.colkey = o; o += (size_t)P * w.n2p * sizeof(unsigned long long);
    w.mutual = o; o += (size_t)P * w.n1p * sizeof(int);
    w.total = o;
    return w;
}

__device__ __forceinline__ unsigned ordered_u32(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ------------------------------------------------------------------------------------------------ match_prep
// One CTA = (cloud, tile of 64 rows), 256 threads, one 16-byte chunk of the image per thread and channel chunk.  Per
// chunk the 16 x 64 block goes through shared memory so that global reads are coalesced along the source's contiguous
// axis and image writes are 16-byte chunks in address order.
constexpr int kPrepRows = 64;
constexpr int kPrepThreads = kPrepRows * 4;
constexpr int kPrepLd = kPrepRows + 2;                       // 4*q*ld mod 32 = {0,8,16,24}: conflict-free chunk reads
constexpr int kPrepStages = 4;                               // chunks in flight per CTA (cp.async groups)

struct PrepSide { const float* desc; float* img; float* nrm; unsigned long long* key; int n, npad; };

__device__ __forceinline__ void cp_async4(float* sdst, const float* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(ri_smem_u32(sdst)), "l"(gsrc) : "memory");
}

// Both descriptor sets in one launch (blockIdx.z = side).  The staging of a chunk is asynchronous (cp.async, one commit
// group per chunk, kPrepStages - 1 chunks in flight): with plain loads every iteration exposed a full memory latency
// (the shared-memory store waits for its load), 32 of them per CTA — 76 us per side where the traffic needs 22.
__global__ void __launch_bounds__(kPrepThreads)
match_prep_kernel(PrepSide s0, PrepSide s1, int C, int Cp, int point_major)
{
    __shared__ float s[kPrepStages][kChunkK * kPrepLd];
    const PrepSide S = blockIdx.z == 0 ? s0 : s1;
    const int n = S.n, npad = S.npad;
    const int cloud = blockIdx.y;
    const int r0 = blockIdx.x * kPrepRows;
    if (r0 >= npad) return;
    const int u = threadIdx.x;                               // 16-byte chunk id inside the (64 rows x 16 ch) block
    const int i_loc = ((u >> 5) << 3) | (u & 7);
    const int q = (u >> 3) & 3;
    const float* D = S.desc + (size_t)cloud * C * n;
    float* I = S.img + (size_t)cloud * npad * Cp;
    const size_t plane = (size_t)npad * kChunkK;             // floats per chunk
    const int nk = Cp / kChunkK;
    double acc = 0.0;

    auto stage = [&](int kc, float* dst) {
        if (kc < nk) {
            if (!point_major) {                              // [C, n]: lanes walk rows (contiguous)
                for (int e = u; e < kChunkK * kPrepRows; e += kPrepThreads) {
                    const int c = e / kPrepRows, i = e % kPrepRows;
                    const int gc = kc * kChunkK + c, gi = r0 + i;
                    if (gc < C && gi < n) cp_async4(dst + c * kPrepLd + i, D + (size_t)gc * n + gi);
                    else dst[c * kPrepLd + i] = 0.f;
                }
            } else {                                         // [n, C]: lanes walk channels (contiguous)
                for (int e = u; e < kChunkK * kPrepRows; e += kPrepThreads) {
                    const int c = e & 15, i = e >> 4;
                    const int gc = kc * kChunkK + c, gi = r0 + i;
                    if (gc < C && gi < n) cp_async4(dst + c * kPrepLd + i, D + (size_t)gi * C + gc);
                    else dst[c * kPrepLd + i] = 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per chunk, empty past the end
    };
#pragma unroll
    for (int p = 0; p < kPrepStages - 1; ++p) stage(p, s[p]);
    for (int kc = 0; kc < nk; ++kc) {
        asm volatile("cp.async.wait_group %0;" :: "n"(kPrepStages - 2) : "memory");   // chunk kc has landed (this thread's part)
        __syncthreads();                                     // ... everyone's part; the buffer of chunk kc - 1 is free again
        stage(kc + kPrepStages - 1, s[(kc + kPrepStages - 1) % kPrepStages]);
        const float* src = s[kc % kPrepStages];
        float4 v;
        v.x = src[(4 * q + 0) * kPrepLd + i_loc]; v.y = src[(4 * q + 1) * kPrepLd + i_loc];
        v.z = src[(4 * q + 2) * kPrepLd + i_loc]; v.w = src[(4 * q + 3) * kPrepLd + i_loc];
        acc = fma((double)v.x, (double)v.x, acc); acc = fma((double)v.y, (double)v.y, acc);
        acc = fma((double)v.z, (double)v.z, acc); acc = fma((double)v.w, (double)v.w, acc);
        // image offset of row (r0 + i_loc), k-core q :  ((row / 8) * 4 + q) * 32 + (row % 8) * 4  floats
        *reinterpret_cast<float4*>(I + (size_t)kc * plane + (size_t)(r0 >> 3) * 128 + (size_t)u * 4) = v;
    }
    // squared norm of row i_loc: the four k-core partials sit in lanes u ^ 8, u ^ 16, u ^ 24
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (q == 0) {
        S.nrm[(size_t)cloud * npad + r0 + i_loc] = (float)acc;
        S.key[(size_t)cloud * npad + r0 + i_loc] = ~0ull;
    }
}

// ------------------------------------------------------------------------------------------------ match_norm
// The pre-pass of the no-image path (channel-major descriptors, indices only): squared norms (fp64 accumulate, rounded once, like
// match_prep) and the reset of the argmin keys — one read of the descriptors at the HBM rate, nothing written but 8 bytes per
// row.  CTA = 32 rows x 8 channel groups; a warp reads 128 contiguous bytes per channel.
__global__ void __launch_bounds__(256)
match_norm_kernel(PrepSide s0, PrepSide s1, int C)
{
    __shared__ double part[8][33];
    const PrepSide S = blockIdx.z == 0 ? s0 : s1;
    const int n = S.n, npad = S.npad;
    const int cloud = blockIdx.y;
    const int r0 = blockIdx.x * 32;
    if (r0 >= npad) return;
    const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
    const int row = r0 + lane;
    const float* D = S.desc + (size_t)cloud * C * n + row;
    double acc = 0.0;
    if (row < n) {
        int c = cg;
        for (; c + 56 < C; c += 64) {                                 // eight loads in flight
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __ldg(D + (size_t)(c + 8 * q) * n);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc = fma((double)v[q], (double)v[q], acc);
        }
        for (; c < C; c += 8) { const float v = __ldg(D + (size_t)c * n); acc = fma((double)v, (double)v, acc); }
    }
    part[cg][lane] = acc;
    __syncthreads();
    if (cg == 0 && row < npad) {
        double t = part[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += part[q][lane];
        S.nrm[(size_t)cloud * npad + row] = (float)t;
        S.key[(size_t)cloud * npad + row] = ~0ull;
    }
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one box of a 3-D tensor map (columns, channels, pair) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
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
// K-major, no swizzle: start address, leading-dimension (K) byte offset, stride-dimension (M/N) byte offset,
// descriptor version 1 (Blackwell) in bits 46-47.  (cute/arch/mma_sm100_desc.hpp::SmemDescriptor)
// (Pinned on a B200: with the two strides exchanged every argmin is wrong.)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, N = 256, M = 128 (cute/arch/mma_sm100_desc.hpp::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);

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

// Column minima of a 32-row x 16-column block held one row per lane (x[c] = this lane's key of column c), by a butterfly:
// at offset 16 / 8 / 4 / 2 a lane keeps the half of its columns selected by that bit of its lane id, hands the other half to
// its partner and takes the minimum with what it receives; the last step folds the lane pair.  Afterwards lanes 2c and 2c + 1
// both hold the minimum of column c.  16 exchanges of 64-bit keys per 16 columns, no warp-wide reduction instructions: the
// redux.sync + ballot per element this replaces made the epilogue (16 us per tile) the kernel's bottleneck.
__device__ __forceinline__ unsigned long long col_min16(unsigned long long (&x)[16], int lane)
{
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool up = (lane & (half << 1)) != 0;          // offsets 16, 8, 4, 2 for half = 8, 4, 2, 1
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const unsigned long long send = up ? x[i] : x[i + half];
            const unsigned long long keep = up ? x[i + half] : x[i];
            const unsigned long long got = __shfl_xor_sync(0xffffffffu, send, half << 1);
            x[i] = got < keep ? got : keep;
        }
    }
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, x[0], 1);
    return other < x[0] ? other : x[0];
}

// ------------------------------------------------------------------------------------------------ match_gemm
struct GemmSmem {                                            // after the stage ring
    unsigned long long colkey[4][kTileN];                    // per epilogue warp, per column
    float n2[kTileN];
    unsigned long long raw[8], full[8], empty[8], acc_full[2], acc_empty[2];      // rings of up to 8 stages
    uint32_t tmem_base;
};

// hi / lo split of 16-byte k-core slots: x -> (x & 0xffffe000), x - hi in the lo plane.
// All loads are issued before the first store (the compiler must assume the stores alias the later loads otherwise,
// which serialises twelve load -> store round trips per stage).
__device__ __forceinline__ void split4(float4 v, float4& h, float4& l)
{
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = __fsub_rn(v.x, h.x);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = __fsub_rn(v.y, h.y);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = __fsub_rn(v.z, h.z);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = __fsub_rn(v.w, h.w);
}

// PERSISTENT: one CTA per SM walks the tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the stage ring runs on across tile
// boundaries and the accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of a tile, the pipeline
// fill of the next one and every per-CTA start-up cost (launch, barrier init, TMEM allocation) overlap the tensor work
// instead of adding to it.  Measured with one CTA per tile: T(C) = 92 us + 0.26 us x C for 32 pairs — at C = 512 the 7 waves of
// tiles spent 40 % of the kernel outside their main loops.
//   warps 0-7            epilogue (TMEM lanes 32 (w % 4) .., columns 128 (w / 4) ..): wait acc_full[buf] -> tcgen05.ld ->
//                        keys -> arrive acc_empty[buf]
//   warps 8 .. 8+4G-1    G converter groups, group g owns the stages with (global stage count) % G == g
//   then                 one TMA producer warp, one MMA issuer / TMEM allocator warp
constexpr int kEpiWarps = 8;                               // two per TMEM lane quarter, 128 columns each
constexpr int kConvWarp0 = kEpiWarps;
constexpr int kPThreads = 32 * (kEpiWarps + 4 * kConvGroups + 2);
static_assert(kStages % kConvGroups == 0, "each stage must belong to one converter group (one phase bit per barrier)");

__global__ void __launch_bounds__(kPThreads, 1)
match_gemm_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                  const float* __restrict__ nrm1, const float* __restrict__ nrm2,
                  int n1, int n2, int n1p, int n2p, int Cp, int tiles_m, int tiles_n, int total_tiles,
                  unsigned long long* __restrict__ rowkey, unsigned long long* __restrict__ colkey,
                  unsigned long long* __restrict__ dbg)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    GemmSmem* S = reinterpret_cast<GemmSmem*>(smem + (size_t)kStages * kStageBytes);
    // debug timeline (tools/exp_match_timeline.py): CTA 0 stamps %globaltimer at a few events of its first 8 tiles
    auto stamp = [&](int it, int ev) {
        if (dbg != nullptr && blockIdx.x == 0 && it < 8) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            dbg[it * 8 + ev] = tt;
        }
    };
    const uint32_t ring = ri_smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const int nk = Cp / kChunkK;
    // stage layout: [A hi 8 KB][A lo 8 KB][B hi 16 KB][B lo 16 KB]; the raw fp32 tiles land in the hi regions
    constexpr int kAHi = 0, kALo = kABytes, kBHi = 2 * kABytes, kBLo = 2 * kABytes + kBBytes;
    constexpr int kProdWarp = kConvWarp0 + 4 * kConvGroups, kMmaWarp = kProdWarp + 1;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto decode = [&](int it, int& pair, int& m0, int& c0) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int per = tiles_m * tiles_n;
        pair = tile / per;
        const int rem = tile - pair * per;
        m0 = (rem / tiles_n) * kTileM;
        c0 = (rem - (rem / tiles_n) * tiles_n) * kTileN;
    };

    if (t == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(ri_smem_u32(&S->raw[s]), 1);           // TMA producer's expect_tx arrival + the bytes
            mbar_init(ri_smem_u32(&S->full[s]), 4);          // one arrival per converter warp of the owning group
            mbar_init(ri_smem_u32(&S->empty[s]), 1);         // tcgen05.commit
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(ri_smem_u32(&S->acc_full[q]), 1);      // tcgen05.commit after a tile's last MMA
            mbar_init(ri_smem_u32(&S->acc_empty[q]), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ri_fence_proxy_async_smem();
    }
    if (warp == kMmaWarp) {                                  // TMEM: 2 x 256 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(ri_smem_u32(&S->tmem_base)), "r"((uint32_t)(2 * kTileN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp == kProdWarp) {
        if (lane == 0) {                                     // ---- TMA producer: raw fp32 tiles, 24 KB per stage
            const size_t planeA = (size_t)n1p * kRowBytes, planeB = (size_t)n2p * kRowBytes;
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                int pair, m0, c0;
                decode(it, pair, m0, c0);
                const uint8_t* A = reinterpret_cast<const uint8_t*>(img1) + (size_t)pair * n1p * Cp * 4;
                const uint8_t* B = reinterpret_cast<const uint8_t*>(img2) + (size_t)pair * n2p * Cp * 4;
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStages;
                    const uint32_t ph = (g / kStages) & 1;
                    mbar_wait(ri_smem_u32(&S->empty[s]), ph ^ 1);
                    const uint32_t bar = ri_smem_u32(&S->raw[s]);
                    mbar_expect_tx(bar, kABytes + kBBytes);
                    const uint32_t dst = ring + s * kStageBytes;
                    bulk_g2s(dst + kAHi, A + (size_t)kc * planeA + (size_t)m0 * kRowBytes, kABytes, bar);
                    bulk_g2s(dst + kBHi, B + (size_t)kc * planeB + (size_t)c0 * kRowBytes, kBBytes, bar);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {                                     // ---- MMA issuer
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                stamp(it, 0);
                mbar_wait(ri_smem_u32(&S->acc_empty[buf]), ((it >> 1) & 1) ^ 1);     // the epilogue has drained this buffer
                tc_fence_after();
                stamp(it, 1);
                const uint32_t acc = tmem + (uint32_t)(buf * kTileN);
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStages;
                    const uint32_t ph = (g / kStages) & 1;
                    mbar_wait(ri_smem_u32(&S->full[s]), ph);
                    tc_fence_after();
                    const uint32_t base = ring + s * kStageBytes;
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks) {
                        const uint32_t koff = ks * 2 * kLBO;     // one K-step = two 16-byte k-cores
                        const uint64_t a_hi = smem_desc(base + kAHi + koff), a_lo = smem_desc(base + kALo + koff);
                        const uint64_t b_hi = smem_desc(base + kBHi + koff), b_lo = smem_desc(base + kBLo + koff);
                        tc_mma_tf32(acc, a_lo, b_hi, kIdesc, (kc | ks) != 0);      // small terms first
                        tc_mma_tf32(acc, a_hi, b_lo, kIdesc, 1);
                        tc_mma_tf32(acc, a_hi, b_hi, kIdesc, 1);
                    }
                    tc_commit(ri_smem_u32(&S->empty[s]));        // stage reusable once these MMAs have read it
                }
                tc_commit(ri_smem_u32(&S->acc_full[buf]));       // this tile's accumulator is complete
                stamp(it, 2);
            }
        }
    } else if (warp >= kConvWarp0) {
        // ---- converters: thread tg of a group splits A-tile row tg and B-tile rows tg, tg + 128 (4 k-core slots each) of
        //      the group's stages.  One stage is a serial chain for its four warps (wait -> 12 LDS -> split -> 12 STS ->
        //      proxy fence -> arrive, ~2.4k clk against 888 clk of MMA work): with a single group the tensor pipe waited
        //      on it (31 % busy); kConvGroups groups convert that many stages at the same time.
        const int ct = t - kConvWarp0 * 32;
        const int grp = ct >> 7, tg = ct & 127;
        const uint32_t slot_a = (uint32_t)(tg >> 3) * kSBO + (uint32_t)(tg & 7) * 16;      // row tg inside a plane
        const uint32_t slot_b1 = (uint32_t)((tg + 128) >> 3) * kSBO + (uint32_t)(tg & 7) * 16;
        const int total = my_tiles * nk;
        for (int g = grp; g < total; g += kConvGroups) {
            const int s = g % kStages;
            const uint32_t ph = (g / kStages) & 1;
            mbar_wait(ri_smem_u32(&S->raw[s]), ph);
            uint8_t* st = smem + (size_t)s * kStageBytes;
            float4 v[12];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[3 * q + 0] = *reinterpret_cast<const float4*>(st + kAHi + slot_a + q * kLBO);
                v[3 * q + 1] = *reinterpret_cast<const float4*>(st + kBHi + slot_a + q * kLBO);
                v[3 * q + 2] = *reinterpret_cast<const float4*>(st + kBHi + slot_b1 + q * kLBO);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 h, l;
                split4(v[3 * q + 0], h, l);
                if (kWriteHi) *reinterpret_cast<float4*>(st + kAHi + slot_a + q * kLBO) = h;
                *reinterpret_cast<float4*>(st + kALo + slot_a + q * kLBO) = l;
                split4(v[3 * q + 1], h, l);
                if (kWriteHi) *reinterpret_cast<float4*>(st + kBHi + slot_a + q * kLBO) = h;
                *reinterpret_cast<float4*>(st + kBLo + slot_a + q * kLBO) = l;
                split4(v[3 * q + 2], h, l);
                if (kWriteHi) *reinterpret_cast<float4*>(st + kBHi + slot_b1 + q * kLBO) = h;
                *reinterpret_cast<float4*>(st + kBLo + slot_b1 + q * kLBO) = l;
            }
            ri_fence_proxy_async_smem();                     // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ri_smem_u32(&S->full[s])) : "memory");
        }
    } else {
        // ---- epilogue warps: warp w reads TMEM lanes [32 (w % 4), +32) = tile rows, columns [128 (w / 4), +128)
        const int wq = warp & 3, half = warp >> 2, tr = t & 127;
        constexpr int kColsPerHalf = kTileN / (kEpiWarps / 4);
        for (int it = 0; it < my_tiles; ++it) {
            int pair, m0, c0;
            decode(it, pair, m0, c0);
            const int buf = it & 1;
            for (int j = t; j < kTileN; j += 32 * kEpiWarps) S->n2[j] = nrm2[(size_t)pair * n2p + c0 + j];
            const int gi = m0 + tr;
            const bool row_ok = gi < n1;
            const float na = nrm1[(size_t)pair * n1p + gi];
            asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");        // n2 staged; previous tile's colkey consumed
            if (t == 0) stamp(it, 3);
            mbar_wait(ri_smem_u32(&S->acc_full[buf]), (it >> 1) & 1);
            tc_fence_after();
            if (t == 0) stamp(it, 4);
            float best = 0.f; int best_j = -1;
            for (int cc = half * kColsPerHalf; cc < (half + 1) * kColsPerHalf; cc += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * kTileN + cc), v);
                // row argmin in registers; column argmin of the 32 rows x 32 columns block by a shuffle butterfly over packed
                // (ordered distance, row) keys, 16 columns at a time (col_min16): lowest row wins ties, as np.argmin
                const unsigned rowid = (unsigned)(m0 + wq * 32 + lane);
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    unsigned long long x[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int ce = hc * 16 + e;
                        const int j = c0 + cc + ce;
                        const float d = __fmaf_rn(-2.0f, __uint_as_float(v[ce]), __fadd_rn(na, S->n2[cc + ce]));
                        if (j < n2 && (best_j < 0 || d < best)) { best = d; best_j = j; }
                        x[e] = ((unsigned long long)(row_ok ? ordered_u32(d) : 0xffffffffu) << 32) | rowid;
                    }
                    const unsigned long long mn = col_min16(x, lane);
                    if ((lane & 1) == 0) S->colkey[wq][cc + hc * 16 + (lane >> 1)] = mn;
                }
            }
            // the accumulator buffer has been read: hand it back to the MMA issuer before the (slow) global atomics
            tc_fence_before();
            __syncwarp();
            if (t == 0) stamp(it, 5);
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ri_smem_u32(&S->acc_empty[buf])) : "memory");
            if (row_ok && best_j >= 0)
                atomicMin(rowkey + (size_t)pair * n1p + gi, ((unsigned long long)ordered_u32(best) << 32) | (unsigned)best_j);
            asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");        // all column keys are in smem
            for (int j = t; j < kTileN; j += 32 * kEpiWarps) {
                if (c0 + j >= n2) continue;
                unsigned long long k0 = S->colkey[0][j];
                const unsigned long long k1 = S->colkey[1][j], k2 = S->colkey[2][j], k3 = S->colkey[3][j];
                k0 = k1 < k0 ? k1 : k0; k0 = k2 < k0 ? k2 : k0; k0 = k3 < k0 ? k3 : k0;
                if ((unsigned)(k0 >> 32) != 0xffffffffu) atomicMin(colkey + (size_t)pair * n2p + c0 + j, k0);
            }
            if (t == 0) stamp(it, 6);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)(2 * kTileN)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ match_gemm_pair
// The persistent kernel on a CTA PAIR (cluster of 2, tcgen05 cta_group::2).  The pair walks 256 x 256 tiles: each CTA owns 128
// rows of f1 (its A tile, its 128 TMEM lanes in both accumulator buffers) and stages HALF of the f2 tile (128 rows of B); the
// leader's single thread issues M256 N256 K8 MMAs and the hardware hands each SM the other half of B.  Per SM and stage the
// shared-memory port then serves 16 KB of TMA writes + 32 KB of converter traffic + 48 KB of operand reads instead of
// 24 + 48 + 72: the single-CTA main loop runs at the shared-memory roofline (1100 clk per stage for 888 clk of MMA work).
//   raw[s]                 per CTA: its own TMA bytes
//   full[s]                in the LEADER: 4 converter warps of each CTA (remote mbarrier.arrive through mapa)
//   empty[s], acc_full[b]  in both CTAs, signalled by tcgen05.commit ... multicast::cluster (mask 0b11)
//   acc_empty[b]           in the LEADER: the 8 epilogue warps of each CTA
constexpr int kPairTileM = 256;
constexpr int kHalfN = kTileN / 2;                           // rows of f2 staged per CTA
constexpr int kB2Bytes = kHalfN * kRowBytes;                 // 8 KB (one plane)
constexpr int kStage2Bytes = 2 * kABytes + 2 * kB2Bytes;     // 32 KB
constexpr int kStagesP = 6;                                  // deeper ring: every stage hand-over crosses the pair twice
constexpr int kConvGroupsP = 3;
constexpr int kPThreadsP = 32 * (kEpiWarps + 4 * kConvGroupsP + 2);
static_assert(kStagesP % kConvGroupsP == 0, "each stage must belong to one converter group");
constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                             ((uint32_t)(kPairTileM >> 4) << 24);

// The no-image path: operands straight from the channel-major descriptors [C, n] by tensor maps.  MN-major tf32 operands have
// ONE legal shared-memory layout, the 128-byte swizzle with 32-byte atomicity (UMMA layout type 1 = SWIZZLE_128B_BASE32B, tensor-
// map swizzle 128B_ATOM_32B: the 32-byte unit index, address bits 5-6, is XORed with the row index, bits 7-8): an atom is 4
// channels x 32 rows of the operand = 4 rows of 128 bytes.  A box of 32 columns x 16 channels lands as 16 such rows = four atoms
// in K; a 128-row operand is four boxes, 2 KB apart.  Descriptor: leading (MN) byte offset 2048 between the 32-row atoms, stride
// (K) byte offset 512 between the 4-channel groups; one K = 8 step is two atoms deep, the second step of a stage starts 1024
// bytes further.  Instruction descriptor: bits 15 / 16 = A / B are MN-major.  (With the 16-byte-atomicity 128B swizzle, type 2,
// the kernel runs and every argmin is wrong — measured.)
constexpr unsigned kMnLBO = 2048, kMnSBO = 512, kMnKStep = 1024;
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(kMnLBO >> 4) << 16) | ((uint64_t)(kMnSBO >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
constexpr uint32_t kIdesc2Mn = kIdesc2 | (1u << 15) | (1u << 16);

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity)      // acquire at cluster scope
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <bool TMA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreadsP, 1)
match_gemm_pair_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                  const float* __restrict__ nrm1, const float* __restrict__ nrm2,
                  int n1, int n2, int n1p, int n2p, int Cp, int tiles_m, int tiles_n, int total_tiles,
                  unsigned long long* __restrict__ rowkey, unsigned long long* __restrict__ colkey,
                  unsigned long long* __restrict__ dbg,
                  const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2, int tma4)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    GemmSmem* S = reinterpret_cast<GemmSmem*>(smem + (size_t)kStagesP * kStage2Bytes);
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;
    // debug timeline (tools/exp_match_timeline.py): CTA 0 stamps %globaltimer at a few events of its first 8 tiles
    auto stamp = [&](int it, int ev) {
        if (dbg != nullptr && blockIdx.x == 0 && it < 8) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
            dbg[it * 8 + ev] = tt;
        }
    };
    const uint32_t ring = ri_smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const int nk = Cp / kChunkK;
    // stage layout: [A hi 8 KB][A lo 8 KB][B hi 16 KB][B lo 16 KB]; the raw fp32 tiles land in the hi regions
    constexpr int kAHi = 0, kALo = kABytes, kBHi = 2 * kABytes, kBLo = 2 * kABytes + kB2Bytes;
    constexpr int kProdWarp = kConvWarp0 + 4 * kConvGroupsP, kMmaWarp = kProdWarp + 1;
    const int my_tiles = (cluster_id < total_tiles) ? (total_tiles - cluster_id + nclusters - 1) / nclusters : 0;
    auto decode = [&](int it, int& pair, int& m0, int& c0) {
        const int tile = cluster_id + it * nclusters;                 // a 256 x 256 tile of the pair
        const int per = tiles_m * tiles_n;
        pair = tile / per;
        const int rem = tile - pair * per;
        m0 = (rem / tiles_n) * kPairTileM + (int)rank * kTileM;      // this CTA's 128 rows of f1
        c0 = (rem - (rem / tiles_n) * tiles_n) * kTileN;
    };

    if (t == 0) {
        for (int s = 0; s < kStagesP; ++s) {
            mbar_init(ri_smem_u32(&S->raw[s]), 1);           // TMA producer's expect_tx arrival + the bytes
            mbar_init(ri_smem_u32(&S->full[s]), 8);          // the owning group's 4 converter warps in each CTA (used in the leader)
            mbar_init(ri_smem_u32(&S->empty[s]), 1);         // tcgen05.commit
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(ri_smem_u32(&S->acc_full[q]), 1);      // tcgen05.commit after a tile's last MMA
            mbar_init(ri_smem_u32(&S->acc_empty[q]), 2 * kEpiWarps);                 // both CTAs' epilogue warps (used in the leader)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ri_fence_proxy_async_smem();
    }
    if (warp == kMmaWarp) {                                  // TMEM: 2 x 256 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(ri_smem_u32(&S->tmem_base)), "r"((uint32_t)(2 * kTileN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // both CTAs' barriers exist before anything crosses the pair
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp == kProdWarp) {
        if (lane == 0) {                                     // ---- TMA producer: raw fp32 tiles, 24 KB per stage
            const size_t planeA = (size_t)n1p * kRowBytes, planeB = (size_t)n2p * kRowBytes;
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                int pair, m0, c0;
                decode(it, pair, m0, c0);
                const uint8_t* A = reinterpret_cast<const uint8_t*>(img1) + (size_t)pair * n1p * Cp * 4;
                const uint8_t* B = reinterpret_cast<const uint8_t*>(img2) + (size_t)pair * n2p * Cp * 4;
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStagesP;
                    const uint32_t ph = (g / kStagesP) & 1;
                    mbar_wait_cl(ri_smem_u32(&S->empty[s]), ph ^ 1);
                    const uint32_t bar = ri_smem_u32(&S->raw[s]);
                    mbar_expect_tx(bar, kABytes + kB2Bytes);
                    const uint32_t dst = ring + s * kStage2Bytes;
                    if constexpr (TMA) {                          // four boxes of 32 rows x 16 channels per operand
                        if (tma4) {                               // ... as one 4-D box each where the tensor allows it
                            tma_load_4d(dst + kAHi, &tmap1, 0, kc * kChunkK, m0 >> 5, pair, bar);
                            tma_load_4d(dst + kBHi, &tmap2, 0, kc * kChunkK, (c0 + (int)rank * kHalfN) >> 5, pair, bar);
                        } else
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            tma_load_3d(dst + kAHi + q * 2048, &tmap1, m0 + 32 * q, kc * kChunkK, pair, bar);
                            tma_load_3d(dst + kBHi + q * 2048, &tmap2, c0 + (int)rank * kHalfN + 32 * q, kc * kChunkK, pair, bar);
                        }
                    } else {
                        bulk_g2s(dst + kAHi, A + (size_t)kc * planeA + (size_t)m0 * kRowBytes, kABytes, bar);
                        bulk_g2s(dst + kBHi, B + (size_t)kc * planeB + (size_t)(c0 + (int)rank * kHalfN) * kRowBytes, kB2Bytes, bar);
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && rank == 0) {                        // ---- MMA issuer: the leader CTA only
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                stamp(it, 0);
                mbar_wait_cl(ri_smem_u32(&S->acc_empty[buf]), ((it >> 1) & 1) ^ 1);     // the epilogue has drained this buffer
                tc_fence_after();
                stamp(it, 1);
                const uint32_t acc = tmem + (uint32_t)(buf * kTileN);
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStagesP;
                    const uint32_t ph = (g / kStagesP) & 1;
                    mbar_wait_cl(ri_smem_u32(&S->full[s]), ph);
                    tc_fence_after();
                    const uint32_t base = ring + s * kStage2Bytes;
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ++ks) {
                        if constexpr (TMA) {
                            const uint32_t koff = ks * kMnKStep; // one K-step = the next 8 channels = two atoms deeper
                            const uint64_t a_hi = smem_desc_mn(base + kAHi + koff), a_lo = smem_desc_mn(base + kALo + koff);
                            const uint64_t b_hi = smem_desc_mn(base + kBHi + koff), b_lo = smem_desc_mn(base + kBLo + koff);
                            tc_mma2_tf32(acc, a_lo, b_hi, kIdesc2Mn, (kc | ks) != 0);
                            tc_mma2_tf32(acc, a_hi, b_lo, kIdesc2Mn, 1);
                            tc_mma2_tf32(acc, a_hi, b_hi, kIdesc2Mn, 1);
                        } else {
                            const uint32_t koff = ks * 2 * kLBO;     // one K-step = two 16-byte k-cores
                            const uint64_t a_hi = smem_desc(base + kAHi + koff), a_lo = smem_desc(base + kALo + koff);
                            const uint64_t b_hi = smem_desc(base + kBHi + koff), b_lo = smem_desc(base + kBLo + koff);
                            tc_mma2_tf32(acc, a_lo, b_hi, kIdesc2, (kc | ks) != 0);    // small terms first
                            tc_mma2_tf32(acc, a_hi, b_lo, kIdesc2, 1);
                            tc_mma2_tf32(acc, a_hi, b_hi, kIdesc2, 1);
                        }
                    }
                    tc_commit2(ri_smem_u32(&S->empty[s]));       // stage reusable in BOTH CTAs once these MMAs have read it
                }
                tc_commit2(ri_smem_u32(&S->acc_full[buf]));      // this tile's accumulators are complete in both CTAs
                stamp(it, 2);
            }
        }
    } else if (warp >= kConvWarp0) {
        // ---- converters: thread tg of a group splits A-tile row tg and B-tile rows tg, tg + 128 (4 k-core slots each) of
        //      the group's stages.  One stage is a serial chain for its four warps (wait -> 12 LDS -> split -> 12 STS ->
        //      proxy fence -> arrive, ~2.4k clk against 888 clk of MMA work): with a single group the tensor pipe waited
        //      on it (31 % busy); kConvGroupsP groups convert that many stages at the same time.
        const int ct = t - kConvWarp0 * 32;
        const int grp = ct >> 7, tg = ct & 127;
        // K-major image: row tg of a plane, its four k-cores 128 bytes apart.  Tensor-map path: the split is element-wise, so
        // any one-to-one walk over the plane's 512 16-byte slots does: slot tg + 128 q.
        const uint32_t slot_a = TMA ? (uint32_t)tg * 16 : (uint32_t)(tg >> 3) * kSBO + (uint32_t)(tg & 7) * 16;
        constexpr uint32_t kSlotStep = TMA ? 128 * 16 : kLBO;
        const uint32_t full_leader = mapa_shared(ri_smem_u32(&S->full[0]), 0);
        const int total = my_tiles * nk;
        for (int g = grp; g < total; g += kConvGroupsP) {
            const int s = g % kStagesP;
            const uint32_t ph = (g / kStagesP) & 1;
            mbar_wait(ri_smem_u32(&S->raw[s]), ph);
            uint8_t* st = smem + (size_t)s * kStage2Bytes;
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[2 * q + 0] = *reinterpret_cast<const float4*>(st + kAHi + slot_a + q * kSlotStep);
                v[2 * q + 1] = *reinterpret_cast<const float4*>(st + kBHi + slot_a + q * kSlotStep);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 h, l;
                split4(v[2 * q + 0], h, l);
                if (kWriteHi) *reinterpret_cast<float4*>(st + kAHi + slot_a + q * kSlotStep) = h;
                *reinterpret_cast<float4*>(st + kALo + slot_a + q * kSlotStep) = l;
                split4(v[2 * q + 1], h, l);
                if (kWriteHi) *reinterpret_cast<float4*>(st + kBHi + slot_a + q * kSlotStep) = h;
                *reinterpret_cast<float4*>(st + kBLo + slot_a + q * kSlotStep) = l;
            }
            ri_fence_proxy_async_smem();                     // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(full_leader + (uint32_t)s * 8u);
        }
    } else {
        // ---- epilogue warps: warp w reads TMEM lanes [32 (w % 4), +32) = tile rows, columns [128 (w / 4), +128)
        const int wq = warp & 3, half = warp >> 2, tr = t & 127;
        const uint32_t acc_empty_leader = mapa_shared(ri_smem_u32(&S->acc_empty[0]), 0);
        constexpr int kColsPerHalf = kTileN / (kEpiWarps / 4);
        for (int it = 0; it < my_tiles; ++it) {
            int pair, m0, c0;
            decode(it, pair, m0, c0);
            const int buf = it & 1;
            for (int j = t; j < kTileN; j += 32 * kEpiWarps) S->n2[j] = nrm2[(size_t)pair * n2p + c0 + j];
            const int gi = m0 + tr;
            const bool row_ok = gi < n1;
            const float na = nrm1[(size_t)pair * n1p + gi];
            asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");        // n2 staged; previous tile's colkey consumed
            if (t == 0) stamp(it, 3);
            mbar_wait_cl(ri_smem_u32(&S->acc_full[buf]), (it >> 1) & 1);
            tc_fence_after();
            if (t == 0) stamp(it, 4);
            float best = 0.f; int best_j = -1;
            for (int cc = half * kColsPerHalf; cc < (half + 1) * kColsPerHalf; cc += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * kTileN + cc), v);
                // row argmin in registers; column argmin of the 32 rows x 32 columns block by a shuffle butterfly over packed
                // (ordered distance, row) keys, 16 columns at a time (col_min16): lowest row wins ties, as np.argmin
                const unsigned rowid = (unsigned)(m0 + wq * 32 + lane);
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    unsigned long long x[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int ce = hc * 16 + e;
                        const int j = c0 + cc + ce;
                        const float d = __fmaf_rn(-2.0f, __uint_as_float(v[ce]), __fadd_rn(na, S->n2[cc + ce]));
                        if (j < n2 && (best_j < 0 || d < best)) { best = d; best_j = j; }
                        x[e] = ((unsigned long long)(row_ok ? ordered_u32(d) : 0xffffffffu) << 32) | rowid;
                    }
                    const unsigned long long mn = col_min16(x, lane);
                    if ((lane & 1) == 0) S->colkey[wq][cc + hc * 16 + (lane >> 1)] = mn;
                }
            }
            // the accumulator buffer has been read: hand it back to the MMA issuer before the (slow) global atomics
            tc_fence_before();
            __syncwarp();
            if (t == 0) stamp(it, 5);
            if (lane == 0) mbar_arrive_remote(acc_empty_leader + (uint32_t)buf * 8u);
            if (row_ok && best_j >= 0)
                atomicMin(rowkey + (size_t)pair * n1p + gi, ((unsigned long long)ordered_u32(best) << 32) | (unsigned)best_j);
            asm volatile("bar.sync 1, %0;" :: "n"(32 * kEpiWarps) : "memory");        // all column keys are in smem
            for (int j = t; j < kTileN; j += 32 * kEpiWarps) {
                if (c0 + j >= n2) continue;
                unsigned long long k0 = S->colkey[0][j];
                const unsigned long long k1 = S->colkey[1][j], k2 = S->colkey[2][j], k3 = S->colkey[3][j];
                k0 = k1 < k0 ? k1 : k0; k0 = k2 < k0 ? k2 : k0; k0 = k3 < k0 ? k3 : k0;
                if ((unsigned)(k0 >> 32) != 0xffffffffu) atomicMin(colkey + (size_t)pair * n2p + c0 + j, k0);
            }
            if (t == 0) stamp(it, 6);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // neither CTA leaves (or frees TMEM) while the other may still use it
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)(2 * kTileN)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ match_finish
// Two kernels.  match_dist: one thread per row of every pair — corr12[i] = argmin_j, the mutual flag, and dist12[i] =
// |f1_i - f2_corr12[i]|^2 accumulated in fp32 from the re-tiled descriptors (per 16-channel chunk a row is 4 x 16 B at a
// 128 B stride).  This is the part with the memory traffic (the two descriptor rows of every match), so it gets the whole
// machine: as one CTA per pair it ran on 32 of the 148 SMs (106 us per 32 pairs).  match_compact: one CTA per pair —
// corr21[j] = argmin_i and the ordered compaction (idx1, idx2)[0..count) of the mutual matches in ascending i (-1 beyond).
constexpr int kDistThreads = 128;
constexpr int kDistLanes = 4;                                // lanes per row: each takes every 4th 16-channel chunk
constexpr int kDistRows = kDistThreads / kDistLanes;
__global__ void __launch_bounds__(kDistThreads)
match_dist_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int n1, int n2, int n1p, int n2p, int Cp,
                  const unsigned long long* __restrict__ rowkey, const unsigned long long* __restrict__ colkey,
                  int* __restrict__ corr12, float* __restrict__ dist12, int* __restrict__ mutual)
{
    // One thread per row left the machine at 10 % occupancy with a 32-step serial load chain per thread (50 us per 32 pairs for
    // 200 MB of traffic); four lanes per row split the chunks, 4x the loads in flight, partial sums combined in a fixed order.
    const int p = blockIdx.y;
    const int sub = threadIdx.x & (kDistLanes - 1);
    const int i = blockIdx.x * kDistRows + (threadIdx.x >> 2);
    const bool live = i < n1;
    const int ic = live ? i : n1 - 1;
    const unsigned long long* RK = rowkey + (size_t)p * n1p;
    const unsigned long long* CK = colkey + (size_t)p * n2p;
    int j = (int)(unsigned)(RK[ic] & 0xffffffffu);
    const bool sane = (unsigned)j < (unsigned)n2;                    // false only if every distance of the row was NaN
    if (!sane) j = 0;
    const size_t plane1 = (size_t)n1p * kChunkK, plane2 = (size_t)n2p * kChunkK;
    const float* a = img1 + (size_t)p * n1p * Cp + (size_t)(ic >> 3) * 128 + (size_t)(ic & 7) * 4;
    const float* b = img2 + (size_t)p * n2p * Cp + (size_t)(j >> 3) * 128 + (size_t)(j & 7) * 4;
    const int nk = Cp / kChunkK;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                             // no cancellation between norms and dot product
    for (int kc = sub; kc < nk; kc += 2 * kDistLanes) {              // two chunks per iteration: 16 independent 16-byte loads
        float4 x[8], y[8];
        const int kc2 = kc + kDistLanes;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            x[q] = __ldg(reinterpret_cast<const float4*>(a + (size_t)kc * plane1 + q * 32));
            y[q] = __ldg(reinterpret_cast<const float4*>(b + (size_t)kc * plane2 + q * 32));
            if (kc2 < nk) {
                x[4 + q] = __ldg(reinterpret_cast<const float4*>(a + (size_t)kc2 * plane1 + q * 32));
                y[4 + q] = __ldg(reinterpret_cast<const float4*>(b + (size_t)kc2 * plane2 + q * 32));
            } else {
                x[4 + q] = make_float4(0.f, 0.f, 0.f, 0.f); y[4 + q] = x[4 + q];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float df = __fsub_rn(x[q].x, y[q].x); acc[0] = __fmaf_rn(df, df, acc[0]);
            df = __fsub_rn(x[q].y, y[q].y); acc[1] = __fmaf_rn(df, df, acc[1]);
            df = __fsub_rn(x[q].z, y[q].z); acc[2] = __fmaf_rn(df, df, acc[2]);
            df = __fsub_rn(x[q].w, y[q].w); acc[3] = __fmaf_rn(df, df, acc[3]);
        }
    }
    float d = __fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3]));
    d = __fadd_rn(d, __shfl_xor_sync(0xffffffffu, d, 1));
    d = __fadd_rn(d, __shfl_xor_sync(0xffffffffu, d, 2));
    if (live && sub == 0) {
        corr12[(size_t)p * n1 + i] = j;
        mutual[(size_t)p * n1p + i] = (sane && (int)(unsigned)(CK[j] & 0xffffffffu) == i) ? 1 : 0;
        dist12[(size_t)p * n1 + i] = d;
    }
}

// dist12 == NULL (the reference's find_correspondence_one_pair returns indices only): nothing to re-evaluate — unpack c1 and the
// mutual flag from the keys, one thread per row (no descriptor traffic at all: ~3 us against 27).
__global__ void __launch_bounds__(256)
match_unpack_kernel(int n1, int n2, int n1p, int n2p, const unsigned long long* __restrict__ rowkey,
                    const unsigned long long* __restrict__ colkey, int* __restrict__ corr12, int* __restrict__ mutual)
{
    const int p = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n1) return;
    int j = (int)(unsigned)(rowkey[(size_t)p * n1p + i] & 0xffffffffu);
    const bool sane = (unsigned)j < (unsigned)n2;                    // false only if every distance of the row was NaN
    if (!sane) j = 0;
    corr12[(size_t)p * n1 + i] = j;
    mutual[(size_t)p * n1p + i] = (sane && (int)(unsigned)(colkey[(size_t)p * n2p + j] & 0xffffffffu) == i) ? 1 : 0;
}

constexpr int kFinThreads = 512;
__global__ void __launch_bounds__(kFinThreads)
match_compact_kernel(int n1, int n2, int n1p, int n2p, const unsigned long long* __restrict__ colkey,
                     const int* __restrict__ corr12, const int* __restrict__ mutual_flag, int* __restrict__ corr21,
                     int* __restrict__ idx1, int* __restrict__ idx2, int* __restrict__ count)
{
    __shared__ int swarp[kFinThreads / 32];
    __shared__ int sbase;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long* CK = colkey + (size_t)p * n2p;
    for (int j = tid; j < n2; j += kFinThreads) corr21[(size_t)p * n2 + j] = (int)(unsigned)(CK[j] & 0xffffffffu);
    if (tid == 0) sbase = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n1; i0 += kFinThreads) {
        const int i = i0 + tid;
        int j = -1, mutual = 0;
        if (i < n1) { j = corr12[(size_t)p * n1 + i]; mutual = mutual_flag[(size_t)p * n1p + i]; }
        const unsigned bal = __ballot_sync(0xffffffffu, mutual);
        const int before = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) swarp[wid] = __popc(bal);
        __syncthreads();
        int off = sbase;
        for (int w = 0; w < wid; ++w) off += swarp[w];
        if (mutual) { idx1[(size_t)p * n1 + off + before] = i; idx2[(size_t)p * n1 + off + before] = j; }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < kFinThreads / 32; ++w) tot += swarp[w]; sbase += tot; }
        __syncthreads();
    }
    const int total = sbase;
    for (int i = total + tid; i < n1; i += kFinThreads) { idx1[(size_t)p * n1 + i] = -1; idx2[(size_t)p * n1 + i] = -1; }
    if (tid == 0) count[p] = total;
}

}  // namespace

// tensor map of channel-major descriptors [P][C][n] (n innermost): box = 32 columns x 16 channels of one pair, 128-byte swizzle
// with 32-byte atomicity (the one layout MN-major tf32 operands accept), out-of-range columns / channels read as zeros (so neither n
// nor C needs padding)
// (the driver entry point is looked up through the runtime, not linked: the library must load on machines without libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn match_encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// 4-D form of the same tensor (n % 32 == 0): (32 columns | C channels | n / 32 column blocks | P pairs) with the column-block
// stride (128 B) BELOW the channel stride (4 n B), so that one box of 32 x 16 x 4 x 1 lands exactly as the four 3-D boxes of an
// operand do, in one instruction instead of four.
static bool match_make_tmap4(CUtensorMap* map, const float* desc, int P, int C, int n)
{
    const EncodeTiledFn cuTensorMapEncodeTiled = match_encode_fn();
    if (cuTensorMapEncodeTiled == nullptr || (n & 31) != 0) return false;
    const cuuint64_t dims[4] = {32u, (cuuint64_t)C, (cuuint64_t)(n / 32), (cuuint64_t)P};
    const cuuint64_t strides[3] = {(cuuint64_t)n * sizeof(float), 128u, (cuuint64_t)C * n * sizeof(float)};
    const cuuint32_t box[4] = {32u, (cuuint32_t)kChunkK, 4u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(desc), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool match_make_tmap(CUtensorMap* map, const float* desc, int P, int C, int n)
{
    const EncodeTiledFn cuTensorMapEncodeTiled = match_encode_fn();
    if (cuTensorMapEncodeTiled == nullptr) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)C, (cuuint64_t)P};
    const cuuint64_t strides[2] = {(cuuint64_t)n * sizeof(float), (cuuint64_t)C * n * sizeof(float)};
    const cuuint32_t box[3] = {32u, (cuuint32_t)kChunkK, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(desc), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" size_t ri_mutual_nn_workspace_bytes(int P, int C, int n1, int n2)
{
    if (P <= 0 || C <= 0 || n1 <= 0 || n2 <= 0) return 16;
    return match_ws_layout(P, C, n1, n2).total + 1024;
}

extern "C" int ri_mutual_nn_tf32x3(const float* desc1, const float* desc2, int P, int C, int n1, int n2, int point_major,
                                   int* corr12, int* corr21, float* dist12, int* idx1, int* idx2, int* count,
                                   void* workspace, size_t workspace_bytes, void* stream)
{
    if (P < 0 || C <= 0 || n1 < 0 || n2 < 0) return RI_ERR_BAD_ARG;
    if (P > 65535) return RI_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 0) return RI_OK;
    if (n1 == 0 || n2 == 0) {                                // nothing can match
        cudaError_t e = cudaMemsetAsync(count, 0, (size_t)P * sizeof(int), st);
        return e == cudaSuccess ? RI_OK : (int)e;
    }
    const MatchWs L = match_ws_layout(P, C, n1, n2);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    if (workspace == nullptr || (size_t)(ws - reinterpret_cast<uint8_t*>(workspace)) + L.total > workspace_bytes)
        return RI_ERR_WORKSPACE;
    float* img1 = reinterpret_cast<float*>(ws + L.img1);
    float* img2 = reinterpret_cast<float*>(ws + L.img2);
    float* nrm1 = reinterpret_cast<float*>(ws + L.nrm1);
    float* nrm2 = reinterpret_cast<float*>(ws + L.nrm2);
    unsigned long long* rowkey = reinterpret_cast<unsigned long long*>(ws + L.rowkey);
    unsigned long long* colkey = reinterpret_cast<unsigned long long*>(ws + L.colkey);

    const PrepSide side1 = {desc1, img1, nrm1, rowkey, n1, L.n1p}, side2 = {desc2, img2, nrm2, colkey, n2, L.n2p};
    const RiEnv& env0 = ri_env();
    // No-image path: indices only (nothing re-reads the descriptors afterwards), channel-major input, CTA-pair GEMM: the operands
    // come straight from the descriptors through tensor maps (MN-major UMMA operands), the pre-pass shrinks to the norms.
    const bool pair_ok = env0.match_pair >= 0 ? env0.match_pair == 1 : (n1 > kTileM && ri_num_sms() >= 2);
    CUtensorMap tm1, tm2;
    bool tma_path = dist12 == nullptr && !point_major && pair_ok && env0.match_tma != 0 && !env0.match_dbg &&
                    (n1 % 4 == 0) && (n2 % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(desc1) | reinterpret_cast<uintptr_t>(desc2)) & 15) == 0;
    int tma4 = 0;
    if (tma_path) {
        tma4 = env0.match_tma != 3 && match_make_tmap4(&tm1, desc1, P, C, n1) && match_make_tmap4(&tm2, desc2, P, C, n2) ? 1 : 0;
        if (!tma4) tma_path = match_make_tmap(&tm1, desc1, P, C, n1) && match_make_tmap(&tm2, desc2, P, C, n2);
    }
    if (tma_path) {
        const int norm_x = (L.n1p > L.n2p ? L.n1p : L.n2p) / 32;
        match_norm_kernel<<<dim3(norm_x, P, 2), 256, 0, st>>>(side1, side2, C);
        RI_LAUNCH_CHECK();
        const size_t smem2 = (size_t)kStagesP * kStage2Bytes + sizeof(GemmSmem) + 1024;
        RI_KERNEL_SETUP(match_gemm_pair_kernel<true>, true, -1);
        const int tiles_n = L.n2p / kTileN;
        const int ptiles_m = (n1 + kPairTileM - 1) / kPairTileM;
        const long long ptotal = (long long)P * ptiles_m * tiles_n;
        if (ptotal > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
        int clusters = ri_num_sms() / 2;
        if (ptotal < clusters) clusters = (int)ptotal;
        match_gemm_pair_kernel<true><<<2 * clusters, kPThreadsP, smem2, st>>>(nullptr, nullptr, nrm1, nrm2, n1, n2, L.n1p, L.n2p, L.Cp,
                                                                              ptiles_m, tiles_n, (int)ptotal, rowkey, colkey, nullptr,
                                                                              tm1, tm2, tma4);
        RI_LAUNCH_CHECK();
        int* mutual_t = reinterpret_cast<int*>(ws + L.mutual);
        match_unpack_kernel<<<dim3((n1 + 255) / 256, P), 256, 0, st>>>(n1, n2, L.n1p, L.n2p, rowkey, colkey, corr12, mutual_t);
        match_compact_kernel<<<P, kFinThreads, 0, st>>>(n1, n2, L.n1p, L.n2p, colkey, corr12, mutual_t, corr21, idx1, idx2, count);
        RI_LAUNCH_CHECK();
        return RI_OK;
    }
    const int prep_x = (L.n1p > L.n2p ? L.n1p : L.n2p) / kPrepRows;
    match_prep_kernel<<<dim3(prep_x, P, 2), kPrepThreads, 0, st>>>(side1, side2, C, L.Cp, point_major);
    RI_LAUNCH_CHECK();

    {
        const size_t smem = (size_t)kStages * kStageBytes + sizeof(GemmSmem) + 1024;
        const RiEnv& env = ri_env();
        const int tiles_m = (n1 + kTileM - 1) / kTileM, tiles_n = L.n2p / kTileN;
        const long long total = (long long)P * tiles_m * tiles_n;
        if (total > 0x7fffffffLL) return RI_ERR_UNSUPPORTED;
        const int grid = total < ri_num_sms() ? (int)total : ri_num_sms();
        unsigned long long* dbg = reinterpret_cast<unsigned long long*>(env.match_dbg ? ws + L.mutual : nullptr);
        // default: CTA pairs (cta_group::2) when a pair tile is not mostly padding; RI_MATCH_PAIR=0 / 1 forces a form
        const bool pair_form = env.match_pair >= 0 ? env.match_pair == 1 : (n1 > kTileM && ri_num_sms() >= 2);
        if (pair_form) {
            // CTA pairs over 256 x 256 tiles (cta_group::2)
            const size_t smem2 = (size_t)kStagesP * kStage2Bytes + sizeof(GemmSmem) + 1024;
            RI_KERNEL_SETUP(match_gemm_pair_kernel<false>, true, -1);
            CUtensorMap none;
            memset(&none, 0, sizeof(none));
            const int ptiles_m = (n1 + kPairTileM - 1) / kPairTileM;
            const long long ptotal = (long long)P * ptiles_m * tiles_n;
            int clusters = ri_num_sms() / 2;
            if (ptotal < clusters) clusters = (int)ptotal;
            match_gemm_pair_kernel<false><<<2 * clusters, kPThreadsP, smem2, st>>>(img1, img2, nrm1, nrm2, n1, n2, L.n1p, L.n2p, L.Cp,
                                                                                 ptiles_m, tiles_n, (int)ptotal, rowkey, colkey, dbg,
                                                                                 none, none, 0);
        } else {
            RI_KERNEL_SETUP(match_gemm_kernel, true, -1);
            match_gemm_kernel<<<grid, kPThreads, smem, st>>>(img1, img2, nrm1, nrm2, n1, n2, L.n1p, L.n2p, L.Cp, tiles_m, tiles_n,
                                                             (int)total, rowkey, colkey, dbg);
        }
        RI_LAUNCH_CHECK();
        if (env.match_dbg) return RI_OK;                     // debug: keep the stamps (match_dist would overwrite them)
    }
    int* mutual = reinterpret_cast<int*>(ws + L.mutual);
    if (dist12 != nullptr)
        match_dist_kernel<<<dim3((n1 + kDistRows - 1) / kDistRows, P), kDistThreads, 0, st>>>(
            img1, img2, n1, n2, L.n1p, L.n2p, L.Cp, rowkey, colkey, corr12, dist12, mutual);
    else
        match_unpack_kernel<<<dim3((n1 + 255) / 256, P), 256, 0, st>>>(n1, n2, L.n1p, L.n2p, rowkey, colkey, corr12, mutual);
    RI_LAUNCH_CHECK();
    match_compact_kernel<<<P, kFinThreads, 0, st>>>(n1, n2, L.n1p, L.n2p, colkey, corr12, mutual, corr21, idx1, idx2, count);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
