// ppf_math.cuh — the point-pair-feature column shared by ppf.cu and the fused k-NN + PPF kernel of knn.cu.
//
// Restates one column of  spherical_ppf_kernel  (/root/reference/PVCNN/modules/functional/src/spherical_ppf/ppf.cu:34-91)
// with the fp32 operation order of its sm_100a PTX (ri_dot3) and the same f64 clamp + acos, so the result is bit-identical.
#pragma once
#include "ri_common.cuh"

struct Ppf4 { float a1, a2, a3, dn; };

// (cx,cy,cz,cn*) = centre and its normal; (x,y,z,n*) = point and its normal.
__device__ __forceinline__ Ppf4 ppf_column(float cx, float cy, float cz, float cnx, float cny, float cnz,
                                           float x, float y, float z, float nx, float ny, float nz)
{
    float dx = __fsub_rn(cx, x), dy = __fsub_rn(cy, y), dz = __fsub_rn(cz, z);              // ppf.cu:53-55
    const float dl = __fsqrt_rn(ri_dot3(dx, dy, dz, dx, dy, dz));
    const float d_norm = __double2float_rn(fmax((double)dl, 1e-20));                        // ppf.cu:56
    dx = __fdiv_rn(dx, d_norm); dy = __fdiv_rn(dy, d_norm); dz = __fdiv_rn(dz, d_norm);
    const float n1 = __fsqrt_rn(ri_dot3(cnx, cny, cnz, cnx, cny, cnz));                     // ppf.cu:61
    const float n2 = __fsqrt_rn(ri_dot3(nx, ny, nz, nx, ny, nz));                           // ppf.cu:62
    Ppf4 o;
    if ((double)n2 <= 1e-10 || (double)n1 <= 1e-10) {                                       // ppf.cu:63-71
        o.a1 = 0.f; o.a2 = 0.f; o.a3 = 0.f; o.dn = 0.f;
        return o;
    }
    cnx = __fdiv_rn(cnx, n1); cny = __fdiv_rn(cny, n1); cnz = __fdiv_rn(cnz, n1);
    nx = __fdiv_rn(nx, n2); ny = __fdiv_rn(ny, n2); nz = __fdiv_rn(nz, n2);
    const double c1 = fmax(fmin((double)ri_dot3(dx, dy, dz, cnx, cny, cnz), 1.0), -1.0);
    const double c2 = fmax(fmin((double)ri_dot3(dx, dy, dz, nx, ny, nz), 1.0), -1.0);
    const double c3 = fmax(fmin((double)ri_dot3(cnx, cny, cnz, nx, ny, nz), 1.0), -1.0);
    o.a1 = __double2float_rn(acos(c1));                                                      // ppf.cu:81-83
    o.a2 = __double2float_rn(acos(c2));
    o.a3 = __double2float_rn(acos(c3));
    o.dn = d_norm;
    return o;
}

