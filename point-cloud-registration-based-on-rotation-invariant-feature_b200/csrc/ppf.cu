// ppf.cu — point-pair features, sm_100a.
//
// Replaces  spherical_ppf_kernel  (/root/reference/PVCNN/modules/functional/src/spherical_ppf/ppf.cu:19-92,
// host wrapper spherical_ppf/ppf.cpp:17-36).  For a (centre, point) column:
//     d = centre - point;  ||d|| = sqrt(fma(dz,dz,fma(dx,dx,dy*dy)));  d_norm = (float)max((double)||d||, 1e-20)
//     d /= d_norm (3 div.rn);  n1 = ||n_centre||, n2 = ||n_point||;  if min(n1,n2) <= 1e-10 -> (0,0,0,0)
//     out = ( acos(clamp(d.n_c)), acos(clamp(d.n_p)), acos(clamp(n_c.n_p)), d_norm )   acos and clamp in f64.
// acos is ill-conditioned at +-1, so the f32 operation order above is reproduced exactly (ri_dot3) and the
// angles go through the same f64 acos; the outputs are then bit-identical to the reference's, not merely 1e-5.
//
// Two entry points:
//   ri_ppf_f32         the reference op: four pre-expanded [B,3,L] operands, one column per thread, grid over
//                      (column tiles, clouds) instead of one CTA per cloud;
//   ri_ppf_gather_f32  the fused form for k-NN neighbourhoods: the cloud (xyz + normal) is staged ONCE per CTA
//                      into shared memory as two float4 per point, the neighbour gather is two LDS.128, the centre
//                      normalisation is done once per centre, and the result is written as [B,4,k,N] (N innermost,
//                      coalesced).  Equals ppf(centres=xyz[:, :, None, :].expand(k), points=gather(xyz, idx), ...).
#include "ri_common.cuh"
#include "ppf_math.cuh"

namespace {

constexpr int kPpfThreads = 256;

// backend argument order: coords (points), center, normals (points), center_normal; all [B,3,L]
__global__ void __launch_bounds__(kPpfThreads)
ppf_columns_kernel(const float* __restrict__ coords, const float* __restrict__ center,
                   const float* __restrict__ normals, const float* __restrict__ center_normal,
                   int L, float* __restrict__ feat)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kPpfThreads + threadIdx.x;
    if (i >= L) return;
    const size_t o3 = (size_t)b * 3 * L;
    const float* P = coords + o3; const float* C = center + o3;
    const float* NP = normals + o3; const float* NC = center_normal + o3;
    const size_t L2 = 2 * (size_t)L;
    const Ppf4 r = ppf_column(C[i], C[i + L], C[i + L2], NC[i], NC[i + L], NC[i + L2],
                              P[i], P[i + L], P[i + L2], NP[i], NP[i + L], NP[i + L2]);
    float* F = feat + (size_t)b * 4 * L;
    F[i] = r.a1; F[i + L] = r.a2; F[i + L2] = r.a3; F[i + L2 + L] = r.dn;
}

// Fused neighbour gather + PPF.  xyz, normals [B,3,N]; idx [B,k,N] (indices into the same cloud);
// out [B,4,k,N].  One CTA = one cloud x one tile of centres; whole cloud staged in smem as float4 pairs.
constexpr int kGatherThreads = 256;

template <bool STAGED>
__global__ void __launch_bounds__(kGatherThreads)
ppf_gather_kernel(const float* __restrict__ xyz, const float* __restrict__ normals, const int* __restrict__ idx,
                  int N, int k, int centres_per_cta, float* __restrict__ out)
{
    extern __shared__ float4 spts[];   // [2*N] when STAGED: (x,y,z,nx), (ny,nz,-,-)
    const int b = blockIdx.y;
    const float* X = xyz + (size_t)b * 3 * N;
    const float* Nn = normals + (size_t)b * 3 * N;
    const size_t N2 = 2 * (size_t)N;
    if (STAGED) {
        for (int t = threadIdx.x; t < N; t += kGatherThreads) {
            spts[2 * t] = make_float4(X[t], X[t + N], X[t + N2], Nn[t]);
            spts[2 * t + 1] = make_float4(Nn[t + N], Nn[t + N2], 0.f, 0.f);
        }
        __syncthreads();
    }
    const int i0 = blockIdx.x * centres_per_cta;
    const int i1 = min(N, i0 + centres_per_cta);
    const int span = i1 - i0;
    const int* I = idx + (size_t)b * k * N;
    float* O = out + (size_t)b * 4 * k * N;
    const size_t kN = (size_t)k * N;
    // work item w = s * span + (i - i0): lanes walk consecutive centres of one neighbour slot s (coalesced)
    for (int w = threadIdx.x; w < span * k; w += kGatherThreads) {
        const int s = w / span;
        const int i = i0 + (w - s * span);
        const int j = I[(size_t)s * N + i];
        float cx, cy, cz, cnx, cny, cnz, x, y, z, nx, ny, nz;
        if (STAGED) {
            const float4 a = spts[2 * i], bq = spts[2 * i + 1];
            cx = a.x; cy = a.y; cz = a.z; cnx = a.w; cny = bq.x; cnz = bq.y;
            const float4 p = spts[2 * j], pq = spts[2 * j + 1];
            x = p.x; y = p.y; z = p.z; nx = p.w; ny = pq.x; nz = pq.y;
        } else {
            cx = X[i]; cy = X[i + N]; cz = X[i + N2]; cnx = Nn[i]; cny = Nn[i + N]; cnz = Nn[i + N2];
            x = X[j]; y = X[j + N]; z = X[j + N2]; nx = Nn[j]; ny = Nn[j + N]; nz = Nn[j + N2];
        }
        const Ppf4 r = ppf_column(cx, cy, cz, cnx, cny, cnz, x, y, z, nx, ny, nz);
        const size_t o = (size_t)s * N + i;
        O[o] = r.a1; O[o + kN] = r.a2; O[o + 2 * kN] = r.a3; O[o + 3 * kN] = r.dn;
    }
}

// The same fused gather + PPF without any shared memory: the cloud comes pre-packed as two float4 per point
// ((x, y, z, nx), (ny, nz, -, -); ri_split_xyz_normals_f32 writes it), so a neighbour is two 16-byte gathers that use
// whole 32-byte sectors and hit L1/L2 (a cloud is 32 KB).  With no shared memory and 40-odd registers the kernel fits
// next to anything — in the front-end step it runs alongside the devoxelizer, whose max-L1 carveout it shares.
__global__ void __launch_bounds__(kGatherThreads)
ppf_gather_packed_kernel(const float4* __restrict__ packed, const int* __restrict__ idx, int N, int k, float* __restrict__ out)
{
    const int b = blockIdx.y;
    const size_t kN = (size_t)k * N;
    const size_t w = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;     // s * N + i : lanes walk consecutive centres
    if (w >= kN) return;
    const int i = (int)(w % N);
    const float4* P = packed + (size_t)b * 2 * N;
    const int j = __ldg(idx + (size_t)b * kN + w);
    const float4 a = __ldg(P + 2 * i), aq = __ldg(P + 2 * i + 1);
    const float4 p = __ldg(P + 2 * j), pq = __ldg(P + 2 * j + 1);
    const Ppf4 r = ppf_column(a.x, a.y, a.z, a.w, aq.x, aq.y, p.x, p.y, p.z, p.w, pq.x, pq.y);
    float* O = out + (size_t)b * 4 * kN + w;
    O[0] = r.a1; O[kN] = r.a2; O[2 * kN] = r.a3; O[3 * kN] = r.dn;
}

}  // namespace

extern "C" int ri_ppf_f32(const float* coords, const float* center, const float* normals,
                          const float* center_normal, int B, int L, float* feat, void* stream)
{
    if (B < 0 || L < 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || L == 0) return RI_OK;
    dim3 grid((L + kPpfThreads - 1) / kPpfThreads, B);
    ppf_columns_kernel<<<grid, kPpfThreads, 0, (cudaStream_t)stream>>>(coords, center, normals, center_normal, L, feat);
    RI_LAUNCH_CHECK();
    return RI_OK;
}

extern "C" int ri_ppf_gather_f32(const float* xyz, const float* normals, const int* idx, int B, int N, int k,
                                 float* out, void* stream)
{
    if (B < 0 || N < 0 || k <= 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || N == 0) return RI_OK;
    // Enough CTAs per cloud to cover the machine a few times over, but not so many that the per-CTA
    // staging of the cloud (24 B/point) dominates.
    const int sms = ri_num_sms();
    int ctas_per_cloud = (4 * sms + B - 1) / B;
    int centres = (N + ctas_per_cloud - 1) / ctas_per_cloud;
    centres = max(32, (centres + 31) / 32 * 32);
    dim3 grid((N + centres - 1) / centres, B);
    const size_t smem = (size_t)N * 2 * sizeof(float4);
    cudaStream_t st = (cudaStream_t)stream;
    if (smem <= 160 * 1024) {
        RI_KERNEL_SETUP(ppf_gather_kernel<true>, true, ri_step_carveout_percent());
        ppf_gather_kernel<true><<<grid, kGatherThreads, smem, st>>>(xyz, normals, idx, N, k, centres, out);
    } else {
        RI_KERNEL_SETUP(ppf_gather_kernel<false>, false, ri_step_carveout_percent());
        ppf_gather_kernel<false><<<grid, kGatherThreads, 0, st>>>(xyz, normals, idx, N, k, centres, out);
    }
    RI_LAUNCH_CHECK();
    return RI_OK;
}

// packed [B,N,8] floats = (x, y, z, nx, ny, nz, 0, 0) per point (16-byte aligned), idx [B,k,N] -> out [B,4,k,N];
// same values as ri_ppf_gather_f32.
extern "C" int ri_ppf_gather_packed_f32(const float* packed, const int* idx, int B, int N, int k, float* out, void* stream)
{
    if (B < 0 || N < 0 || k <= 0 || ((uintptr_t)packed & 15) != 0) return RI_ERR_BAD_ARG;
    if (B > 65535) return RI_ERR_UNSUPPORTED;
    if (B == 0 || N == 0) return RI_OK;
    const size_t kN = (size_t)k * N;
    dim3 grid((unsigned)((kN + kGatherThreads - 1) / kGatherThreads), B);
    // no shared memory, but the carveout PREFERENCE still decides which kernels an SM can host at the same time: ask for
    // the step's split so it runs next to the streaming devoxelizer / grid writer / k-NN (RI_PPF_MAXL1=1: the default split)
    if (!ri_env().ppf_maxl1) RI_KERNEL_SETUP(ppf_gather_packed_kernel, false, ri_step_carveout_percent());
    ppf_gather_packed_kernel<<<grid, kGatherThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(packed), idx, N, k, out);
    RI_LAUNCH_CHECK();
    return RI_OK;
}
