/*
 * ri_oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product package never imports anything under oracle/.
 *
 * Every function restates one reference CUDA kernel (paths relative to /root/reference) in scalar C,
 * one cloud after another, one point after another.  The fp32 arithmetic is written with explicit
 * fmaf()/sqrtf() and explicit double casts in exactly the contraction order nvcc chose for the
 * reference when compiled for sm_100a (read from the PTX; see DESIGN.md "arithmetic pins"), and the file
 * must be compiled with  -ffp-contract=off -fno-fast-math  so that gcc adds no contraction of its own.
 *
 * Parity pin: tests/golden/ holds outputs of the reference's own kernels (oracle/_ref, built from the
 * unmodified sources) run on a B200; tests/test_oracle_golden.py checks this file against them.
 * Known, documented limit: acosf()/atanf() here are glibc's, the reference uses CUDA libdevice's
 * (rsqrt.approx / rcp.approx based).  They differ by <= 1-2 ulp on rare arguments, which can move a point
 * that sits on a spherical cell boundary into the neighbouring cell.  Integer-only paths (KNN order,
 * cube voxel indices, counts, devox corner indices) are bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RI_PI 3.14159265358979311600e+00 /* == acos(-1.0) as evaluated on device: 0x400921FB54442D18 */

/* ------------------------------------------------------------------------------------------------
 * KNN, one direction.  PVCNN/modules/functional/src/knn/knn.cu:5-49 (KnnKernel) with the slot
 * initialisation of knn/knn.cpp:14-17 (dist = 10000.0f, idx = 0).
 * xyz1 [B,c,n] queries, xyz2 [B,c,m] references -> dist [B,k,n], idx [B,k,n].
 * d = fma(d_p,d_p,acc) over channels p = 0..c-1 starting from acc = 0  (PTX: sub.f32 + fma.rn.f32 chain).
 * -----------------------------------------------------------------------------------------------*/
void ri_oracle_knn(const float *xyz1, const float *xyz2, int B, int c, int n, int m, int k,
                   float *dist, int *idx)
{
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *q = xyz1 + (size_t)b * c * n;
        const float *r = xyz2 + (size_t)b * c * m;
        float *res = dist + (size_t)b * k * n;
        int *resi = idx + (size_t)b * k * n;
        for (int i = 0; i < n; ++i) {
            for (int s = 0; s < k; ++s) { res[i + (size_t)s * n] = 10000.0f; resi[i + (size_t)s * n] = 0; }
            for (int j = 0; j < m; ++j) {
                float d = 0.0f;
                for (int p = 0; p < c; ++p) {
                    float df = q[i + (size_t)p * n] - r[j + (size_t)p * m];
                    d = fmaf(df, df, d);
                }
                /* knn.cu:27-31 : strict '<' against the last slot */
                if (d < res[i + (size_t)(k - 1) * n]) {
                    res[i + (size_t)(k - 1) * n] = d;
                    resi[i + (size_t)(k - 1) * n] = j;
                }
                /* knn.cu:33-45 : one bubble pass, strict '<' => stable, lower index wins ties */
                for (int s = k - 1; s > 0; --s) {
                    float a = res[i + (size_t)s * n], bb = res[i + (size_t)(s - 1) * n];
                    if (a < bb) {
                        res[i + (size_t)s * n] = bb; res[i + (size_t)(s - 1) * n] = a;
                        int t = resi[i + (size_t)s * n];
                        resi[i + (size_t)s * n] = resi[i + (size_t)(s - 1) * n];
                        resi[i + (size_t)(s - 1) * n] = t;
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * KNN backward, one direction.  knn/knn.cu:52-78 (KnnGradKernel).  Accumulates into grad1/grad2
 * (caller zero-fills, knn.cpp:44-45).  Sequential accumulation order (the reference uses float atomics).
 * -----------------------------------------------------------------------------------------------*/
void ri_oracle_knn_grad(const float *xyz1, const float *xyz2, const float *gdist, const int *idx,
                        int B, int c, int n, int m, int k, float *grad1, float *grad2)
{
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i)
            for (int s = 0; s < k; ++s) {
                float g = gdist[(size_t)b * k * n + (size_t)s * n + i] * 2.0f;
                if (g >= 20000.0f) continue;                          /* knn.cu:68 */
                int id = idx[(size_t)b * k * n + (size_t)s * n + i];
                for (int p = 0; p < c; ++p) {
                    float a = xyz1[(size_t)b * c * n + (size_t)p * n + i];
                    float bb = xyz2[(size_t)b * c * m + (size_t)p * m + id];
                    float t = g * (a - bb);
                    grad1[(size_t)b * c * n + (size_t)p * n + i] += t;
                    grad2[(size_t)b * c * m + (size_t)p * m + id] += -t;
                }
            }
}

/* ------------------------------------------------------------------------------------------------
 * PPF.  spherical_ppf/ppf.cu:19-92, backend argument order (coords, center, normals, center_normal);
 * the Python wrapper functional/ppf.py:22 swaps (centers, points) into that order.
 * All four inputs [B,3,L], feat [B,4,L] (zero-filled by spherical_ppf/ppf.cpp:29-30).
 * -----------------------------------------------------------------------------------------------*/
static float ri_dot3(float ax, float ay, float az, float bx, float by, float bz)
{   /* nvcc: a.x*b.x + a.y*b.y + a.z*b.z  ->  fma(az,bz, fma(ax,bx, ay*by)) */
    return fmaf(az, bz, fmaf(ax, bx, ay * by));
}

void ri_oracle_ppf(const float *coords, const float *center, const float *normals,
                   const float *center_normal, int B, int L, float *feat)
{
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *P = coords + (size_t)b * 3 * L, *Cc = center + (size_t)b * 3 * L;
        const float *Np = normals + (size_t)b * 3 * L, *Nc = center_normal + (size_t)b * 3 * L;
        float *F = feat + (size_t)b * 4 * L;
        for (int i = 0; i < L; ++i) {
            float x = P[i], y = P[i + L], z = P[i + 2 * (size_t)L];
            float nx = Np[i], ny = Np[i + L], nz = Np[i + 2 * (size_t)L];
            float cx = Cc[i], cy = Cc[i + L], cz = Cc[i + 2 * (size_t)L];
            float cnx = Nc[i], cny = Nc[i + L], cnz = Nc[i + 2 * (size_t)L];
            float dx = cx - x, dy = cy - y, dz = cz - z;                       /* ppf.cu:53-55 */
            float dn = sqrtf(ri_dot3(dx, dy, dz, dx, dy, dz));
            float d_norm = (float)fmax((double)dn, 1e-20);                     /* ppf.cu:56 (double max) */
            dx = dx / d_norm; dy = dy / d_norm; dz = dz / d_norm;
            float n1 = sqrtf(ri_dot3(cnx, cny, cnz, cnx, cny, cnz));           /* ppf.cu:61 */
            float n2 = sqrtf(ri_dot3(nx, ny, nz, nx, ny, nz));                 /* ppf.cu:62 */
            if ((double)n2 <= 1e-10 || (double)n1 <= 1e-10) {                  /* ppf.cu:63-71 */
                F[i] = 0; F[i + L] = 0; F[i + 2 * (size_t)L] = 0; F[i + 3 * (size_t)L] = 0;
                continue;
            }
            cnx = cnx / n1; cny = cny / n1; cnz = cnz / n1;
            nx = nx / n2; ny = ny / n2; nz = nz / n2;
            double a1 = fmax(fmin((double)ri_dot3(dx, dy, dz, cnx, cny, cnz), 1.0), -1.0);
            double a2 = fmax(fmin((double)ri_dot3(dx, dy, dz, nx, ny, nz), 1.0), -1.0);
            double a3 = fmax(fmin((double)ri_dot3(cnx, cny, cnz, nx, ny, nz), 1.0), -1.0);
            F[i] = (float)acos(a1);                                            /* ppf.cu:81-83, f64 acos */
            F[i + L] = (float)acos(a2);
            F[i + 2 * (size_t)L] = (float)acos(a3);
            F[i + 3 * (size_t)L] = d_norm;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Spherical coordinates of one point, shared by binning and spherical devox.
 * spherical_voxelization/spherical_vox.cu:34-56 == interpolate/spherical_trilinear_devox.cu:48-65.
 * Returns 0 if the point is "undefined".
 * -----------------------------------------------------------------------------------------------*/
static int ri_sph_coords(float x, float y, float z, int r, float *gama_o, float *alpha_o, float *beta_o)
{
    float gama = sqrtf(fmaf(z, z, fmaf(x, x, y * y)));
    if (gama == 0.0f || gama >= 1.0f) return 0;
    float t = z / gama;
    if (t > 1.0f || t < -1.0f) return 0;
    float beta = acosf(t);
    if ((double)beta >= RI_PI) return 0;
    float alpha;
    if (x == 0.0f && y != 0.0f) alpha = (float)(((double)(y / fabsf(y)) * RI_PI) * 0.5);
    else if (x == 0.0f && y == 0.0f) alpha = 0.0f;
    else alpha = (float)fma(RI_PI * (double)(1.0f - x / fabsf(x)), 0.5, (double)atanf(y / x));
    alpha = (float)(RI_PI / (double)r + (double)alpha);                        /* :55 */
    if (alpha < 0.0f) alpha = (float)fma(RI_PI, 2.0, (double)alpha);           /* :56 */
    *gama_o = gama; *alpha_o = alpha; *beta_o = beta;
    return 1;
}

/* Spherical binning: spherical_vox.cu:19-77 (spherical_grid_stats_kernel). coords [B,3,N] -> ind [B,N], cnt [B,s] */
void ri_oracle_sph_grid_stats(const float *coords, int B, int N, int r, int *ind, int *cnt)
{
    int r2 = r * r, s = r2 * r;
    float rf = (float)r;
    memset(cnt, 0, sizeof(int) * (size_t)B * s);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *X = coords + (size_t)b * 3 * N;
        for (int i = 0; i < N; ++i) {
            float g, a, be;
            if (!ri_sph_coords(X[i], X[i + N], X[i + 2 * (size_t)N], r, &g, &a, &be)) {
                ind[(size_t)b * N + i] = -1;
                continue;
            }
            int gx = (int)floorf(g * rf);                                      /* :59 */
            int gy = (int)floor((double)((a * rf) * 0.5f) / RI_PI);            /* :60 */
            int gz = (int)floor((double)(be * rf) / RI_PI);                    /* :61 */
            if (gx >= r) gx = r - 1;
            if (gy >= r) gy = r - 1;
            if (gz >= r) gz = r - 1;
            int id = gx * r2 + gy * r + gz;
            ind[(size_t)b * N + i] = id;
            cnt[(size_t)b * s + id] += 1;
        }
    }
}

/* Continuous spherical grid coordinates (diagnostic for the libm-vs-libdevice boundary cases). */
void ri_oracle_sph_grid_cont(const float *coords, int B, int N, int r, double *gc)
{
    float rf = (float)r;
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *X = coords + (size_t)b * 3 * N;
        for (int i = 0; i < N; ++i) {
            float g, a, be;
            double *o = gc + ((size_t)b * N + i) * 3;
            if (!ri_sph_coords(X[i], X[i + N], X[i + 2 * (size_t)N], r, &g, &a, &be)) { o[0] = o[1] = o[2] = -1; continue; }
            o[0] = (double)(g * rf);
            o[1] = (double)((a * rf) * 0.5f) / RI_PI;
            o[2] = (double)(be * rf) / RI_PI;
        }
    }
}

/* Cube binning: voxelization/vox.cu:18-35 (grid_stats_kernel). coords int [B,3,N] */
void ri_oracle_cube_grid_stats(const int *coords, int B, int N, int r, int *ind, int *cnt)
{
    int r2 = r * r, s = r2 * r;
    memset(cnt, 0, sizeof(int) * (size_t)B * s);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const int *X = coords + (size_t)b * 3 * N;
        for (int i = 0; i < N; ++i) {
            int id = X[i] * r2 + X[i + N] * r + X[i + 2 * (size_t)N];           /* vox.cu:31 */
            ind[(size_t)b * N + i] = id;
            cnt[(size_t)b * s + id] += 1;
        }
    }
}

/* Scatter-mean: vox.cu:49-73 == spherical_vox.cu:91-125.  out[c,pos] = sum_i feat[c,i] * (1.0f/cnt[pos]),
 * here summed in ascending point order (the reference order is whatever the float atomics produce). */
void ri_oracle_avg_voxelize(const float *feat, const int *ind, const int *cnt, int B, int C, int N, int s,
                            float *out)
{
    memset(out, 0, sizeof(float) * (size_t)B * C * s);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            int pos = ind[(size_t)b * N + i];
            if (pos == -1) continue;
            int cc = cnt[(size_t)b * s + pos];
            if (cc <= 0) continue;
            float inv = 1.0f / (float)cc;
            for (int c = 0; c < C; ++c)
                out[((size_t)b * C + c) * s + pos] += feat[((size_t)b * C + c) * N + i] * inv;
        }
}

/* Scatter-mean backward: vox.cu:87-111 == spherical_vox.cu:139-163. grad_x[c,i] = grad_y[c,pos] * (1/cnt) */
void ri_oracle_avg_voxelize_grad(const float *grad_y, const int *ind, const int *cnt, int B, int C, int N, int s,
                                 float *grad_x)
{
    memset(grad_x, 0, sizeof(float) * (size_t)B * C * N);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            int pos = ind[(size_t)b * N + i];
            if (pos == -1) continue;
            int cc = cnt[(size_t)b * s + pos];
            if (cc <= 0) continue;
            float inv = 1.0f / (float)cc;
            for (int c = 0; c < C; ++c)
                grad_x[((size_t)b * C + c) * N + i] = 0.0f + grad_y[((size_t)b * C + c) * s + pos] * inv;
        }
}

/* 8-term interpolation shared by both devoxelizers: chain order 001(mul),000,010,011,100,101,110,111 */
static void ri_devox_point(const float *feat, int C, int s, int N, int i, const int id[8], const float w[8], float *outs)
{
    for (int c = 0; c < C; ++c) {
        const float *f = feat + (size_t)c * s;
        float acc = w[1] * f[id[1]];
        acc = fmaf(w[0], f[id[0]], acc);
        acc = fmaf(w[2], f[id[2]], acc);
        acc = fmaf(w[3], f[id[3]], acc);
        acc = fmaf(w[4], f[id[4]], acc);
        acc = fmaf(w[5], f[id[5]], acc);
        acc = fmaf(w[6], f[id[6]], acc);
        acc = fmaf(w[7], f[id[7]], acc);
        outs[(size_t)c * N + i] = acc;
    }
}

static void ri_corner_table(float d1a, float d1b, float d1c, int lo_a, int lo_b, int lo_c, int r, int id[8], float w[8])
{
    /* interpolate/trilinear_devox.cu:46-76: weights (da*db)*dc, hi offsets only where the residual is > 0 */
    float d0a = 1.0f - d1a, d0b = 1.0f - d1b, d0c = 1.0f - d1c;
    w[0] = (d0a * d0b) * d0c; w[1] = (d0a * d0b) * d1c;
    w[2] = (d0a * d1b) * d0c; w[3] = (d0a * d1b) * d1c;
    w[4] = (d1a * d0b) * d0c; w[5] = (d1a * d0b) * d1c;
    w[6] = (d1a * d1b) * d0c; w[7] = (d1a * d1b) * d1c;
    int r2 = r * r;
    int ha = (d1a > 0) ? r2 : 0, hb = (d1b > 0) ? r : 0, hc = (d1c > 0) ? 1 : 0;
    id[0] = lo_a * r2 + lo_b * r + lo_c;
    id[1] = id[0] + hc;
    id[2] = id[0] + hb;
    id[3] = id[2] + hc;
    id[4] = id[0] + ha;
    id[5] = id[4] + hc;
    id[6] = id[4] + hb;
    id[7] = id[6] + hc;
}

/* Cube trilinear devoxelize: interpolate/trilinear_devox.cu:22-106. coords [B,3,N] (grid units, [0,r-1]),
 * feat [B,C,s] -> outs [B,C,N], inds [B,8,N], wgts [B,8,N] */
void ri_oracle_trilinear_devox(const float *coords, const float *feat, int B, int C, int N, int r,
                               float *outs, int *inds, float *wgts)
{
    int s = r * r * r;
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *X = coords + (size_t)b * 3 * N;
        for (int i = 0; i < N; ++i) {
            float x = X[i], y = X[i + N], z = X[i + 2 * (size_t)N];
            float xl = floorf(x), yl = floorf(y), zl = floorf(z);
            int id[8]; float w[8];
            ri_corner_table(x - xl, y - yl, z - zl, (int)xl, (int)yl, (int)zl, r, id, w);
            for (int q = 0; q < 8; ++q) {
                inds[((size_t)b * 8 + q) * N + i] = id[q];
                wgts[((size_t)b * 8 + q) * N + i] = w[q];
            }
            ri_devox_point(feat + (size_t)b * C * s, C, s, N, i, id, w, outs + (size_t)b * C * N);
        }
    }
}

/* Spherical "trilinear" devoxelize with the reference's index quirks:
 * interpolate/spherical_trilinear_devox.cu:23-136.  outs/inds/wgts are zero-filled first
 * (spherical_trilinear_devox.cpp:33-41); undefined points leave their row at zero except inds[0,i] = -1. */
void ri_oracle_sph_trilinear_devox(const float *coords, const float *feat, const int *g_inds,
                                   int B, int C, int N, int r, float *outs, int *inds, float *wgts)
{
    int r2 = r * r, s = r2 * r;
    memset(outs, 0, sizeof(float) * (size_t)B * C * N);
    memset(inds, 0, sizeof(int) * (size_t)B * 8 * N);
    memset(wgts, 0, sizeof(float) * (size_t)B * 8 * N);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *X = coords + (size_t)b * 3 * N;
        for (int i = 0; i < N; ++i) {
            int pos = g_inds[(size_t)b * N + i];
            if (pos == -1) { inds[((size_t)b * 8) * N + i] = -1; continue; }   /* :42-47 */
            float g, a, be;
            if (!ri_sph_coords(X[i], X[i + N], X[i + 2 * (size_t)N], r, &g, &a, &be)) continue;  /* :54,:59 */
            int gg = pos / r2;
            int ga = (pos - gg * r2) / r;
            int gb = pos - gg * r2 - ga * r;
            float g_lo = (float)(gg / r);                                      /* :71 integer division */
            float a_lo = (float)(((RI_PI * 2.0) * (double)ga) / (double)r);    /* :72 radians */
            float b_lo = (float)((RI_PI * (double)gb) / (double)r);            /* :73 radians */
            int id[8]; float w[8];
            ri_corner_table(g - g_lo, a - a_lo, be - b_lo, (int)g_lo, (int)a_lo, (int)b_lo, r, id, w);
            for (int q = 0; q < 8; ++q) {
                inds[((size_t)b * 8 + q) * N + i] = id[q];
                wgts[((size_t)b * 8 + q) * N + i] = w[q];
            }
            ri_devox_point(feat + (size_t)b * C * s, C, s, N, i, id, w, outs + (size_t)b * C * N);
        }
    }
}

/* Devox backward (both variants): trilinear_devox.cu:120-163 / spherical_trilinear_devox.cu:150-194.
 * skip_undefined != 0 reproduces the spherical kernel's `inds[0,i] == -1 -> continue`. */
void ri_oracle_devox_grad(const float *grad_y, const int *inds, const float *wgts, int B, int C, int N, int s,
                          int skip_undefined, float *grad_x)
{
    memset(grad_x, 0, sizeof(float) * (size_t)B * C * s);
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            if (skip_undefined && inds[((size_t)b * 8) * N + i] == -1) continue;
            for (int c = 0; c < C; ++c) {
                float g = grad_y[((size_t)b * C + c) * N + i];
                for (int q = 0; q < 8; ++q)
                    grad_x[((size_t)b * C + c) * s + inds[((size_t)b * 8 + q) * N + i]] += wgts[((size_t)b * 8 + q) * N + i] * g;
            }
        }
}

/* DGCNN voxel-neighbour edge features: PVCNN/modules/pvconv.py:68-90.
 * avg [B,C,s], feat [B,C,N], inds [B,N] -> out [B,2C,N] = cat(feat - avg[:, inds] (0 where inds==-1), feat) */
void ri_oracle_voxel_edge_gather(const float *avg, const float *feat, const int *inds, int B, int C, int N, int s,
                                 float *out)
{
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int i = 0; i < N; ++i) {
                int id = inds[(size_t)b * N + i];
                float f = feat[((size_t)b * C + c) * N + i];
                float rel = (id == -1) ? 0.0f : f - avg[((size_t)b * C + c) * s + id];
                out[((size_t)b * 2 * C + c) * N + i] = rel;
                out[((size_t)b * 2 * C + C + c) * N + i] = f;
            }
}

/* ------------------------------------------------------------------------------------------------
 * Ball query.  PVCNN/modules/functional/src/ball_query/ball_query.cu:19-50 (+ torch::zeros output and
 * r2 = radius * radius in float, ball_query.cpp:18-24).  centers [B,3,M], points [B,3,N] -> idx [B,M,U].
 * d2 = fma(dz,dz, fma(dy,dy, dx*dx)) is the contraction in the reference's sm_100a SASS (FMUL dx, FFMA dy, FFMA dz);
 * the lower bound is the double pow(10,-5) == 1e-5 for every float d2 (no float lies within a double ulp of it).
 * -----------------------------------------------------------------------------------------------*/
void ri_oracle_ball_query(const float *centers, const float *points, int B, int N, int M, float radius, int U, int *idx)
{
    const float r2 = radius * radius;
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b) {
        const float *c = centers + (size_t)b * 3 * M;
        const float *p = points + (size_t)b * 3 * N;
        int *o = idx + (size_t)b * M * U;
        for (int j = 0; j < M; ++j) {
            for (int v = 0; v < U; ++v) o[(size_t)j * U + v] = 0;
            int cnt = 0;
            for (int k = 0; k < N && cnt < U; ++k) {
                float dx = c[j] - p[k], dy = c[j + M] - p[k + N], dz = c[j + 2 * (size_t)M] - p[k + 2 * (size_t)N];
                float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d2 < r2 && (double)d2 > 1e-5) {
                    if (cnt == 0)
                        for (int v = 0; v < U; ++v) o[(size_t)j * U + v] = k;
                    o[(size_t)j * U + cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* Grouping forward / backward.  PVCNN/modules/functional/src/grouping/grouping.cu:18-44, 58-84.
 * feat [B,C,N], idx [B,M,U] -> out [B,C,M,U];  grad_x [B,C,N] = scatter-add of grad_y (summed here in index order). */
void ri_oracle_grouping(const float *feat, const int *idx, int B, int C, int N, int M, int U, float *out)
{
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < C; ++l)
            for (size_t e = 0; e < (size_t)M * U; ++e)
                out[((size_t)b * C + l) * M * U + e] = feat[((size_t)b * C + l) * N + idx[(size_t)b * M * U + e]];
}

void ri_oracle_grouping_grad(const float *grad_y, const int *idx, int B, int C, int N, int M, int U, float *grad_x)
{
    memset(grad_x, 0, (size_t)B * C * N * sizeof(float));
    _Pragma("omp parallel for schedule(dynamic)")
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < C; ++l)
            for (size_t e = 0; e < (size_t)M * U; ++e)
                grad_x[((size_t)b * C + l) * N + idx[(size_t)b * M * U + e]] += grad_y[((size_t)b * C + l) * M * U + e];
}

/* ------------------------------------------------------------------------------------------------
 * Barycentre grid subsampling.  cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:4-106
 * (SampledData: grid_subsampling.h:9-84; PointXYZ float arithmetic: cpp_wrappers/cpp_utils/cloud/cloud.h:40-155).
 * points [N,3], features [N,fdim] or NULL, labels [N,ldim] or NULL -> out_* for M cells, returns M.
 * The reference emits cells in unordered_map iteration order (unspecified); here: ASCENDING cell index.
 * Label ties: the reference returns the first maximum in the inner unordered_map's iteration order
 * (unspecified); here: the smallest label.  keys_out (optional, [N]) receives each output cell's index.
 * -----------------------------------------------------------------------------------------------*/
typedef struct { uint64_t key; int idx; } ri_gs_pair;
static int ri_gs_cmp(const void *a, const void *b)
{
    const ri_gs_pair *x = (const ri_gs_pair *)a, *y = (const ri_gs_pair *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);          /* original point order inside a cell */
}

int ri_oracle_grid_subsample(const float *pts, const float *feats, const int *labels, int N, int fdim, int ldim,
                             float dl, float *out_pts, float *out_feats, int *out_labels, uint64_t *keys_out)
{
    if (N <= 0) return 0;
    float mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
    for (int i = 1; i < N; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = pts[3 * (size_t)i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    const float inv = 1 / dl;                                          /* :27  (1/sampleDl) */
    float org[3];
    for (int a = 0; a < 3; ++a) org[a] = floorf(mn[a] * inv) * dl;     /* :27 */
    const uint64_t nx = (uint64_t)floorf((mx[0] - org[0]) / dl) + 1;   /* :30 */
    const uint64_t ny = (uint64_t)floorf((mx[1] - org[1]) / dl) + 1;   /* :31 */
    ri_gs_pair *pr = (ri_gs_pair *)malloc((size_t)N * sizeof(ri_gs_pair));
    for (int i = 0; i < N; ++i) {
        const uint64_t ix = (uint64_t)floorf((pts[3 * (size_t)i + 0] - org[0]) / dl);     /* :53-56 */
        const uint64_t iy = (uint64_t)floorf((pts[3 * (size_t)i + 1] - org[1]) / dl);
        const uint64_t iz = (uint64_t)floorf((pts[3 * (size_t)i + 2] - org[2]) / dl);
        pr[i].key = ix + nx * iy + nx * ny * iz;
        pr[i].idx = i;
    }
    qsort(pr, (size_t)N, sizeof(ri_gs_pair), ri_gs_cmp);
    int M = 0;
    for (int s0 = 0; s0 < N;) {
        int s1 = s0;
        while (s1 < N && pr[s1].key == pr[s0].key) ++s1;
        const int cnt = s1 - s0;
        for (int a = 0; a < 3; ++a) {
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc = acc + pts[3 * (size_t)pr[s].idx + a];     /* point += p, original order */
            out_pts[3 * (size_t)M + a] = acc * (float)(1.0 / cnt);                         /* :84 point * (1.0 / count) */
        }
        for (int f = 0; f < fdim; ++f) {
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc = acc + feats[(size_t)pr[s].idx * fdim + f];
            out_feats[(size_t)M * fdim + f] = acc / (float)cnt;                            /* :88-92 */
        }
        for (int l = 0; l < ldim; ++l) {
            int best = 0, best_n = 0;
            for (int s = s0; s < s1; ++s) {
                const int v = labels[(size_t)pr[s].idx * ldim + l];
                int n = 0;
                for (int t = s0; t < s1; ++t) n += labels[(size_t)pr[t].idx * ldim + l] == v;
                if (n > best_n || (n == best_n && v < best)) { best = v; best_n = n; }
            }
            out_labels[(size_t)M * ldim + l] = best;                                       /* :96-100 */
        }
        if (keys_out) keys_out[M] = pr[s0].key;
        ++M;
        s0 = s1;
    }
    free(pr);
    return M;
}
