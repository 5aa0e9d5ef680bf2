// gridsub_ref_shim.cpp — C entry point around the reference's OWN grid_subsampling() so that tests can call the
// unmodified reference code through ctypes.  TEST INFRASTRUCTURE ONLY; this file contains no reference code: it is
// compiled together with /root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp and
// /root/reference/cpp_wrappers/cpp_utils/cloud/cloud.cpp, where they lie, by oracle/build_ref.py::build_gridsub()
// into oracle/_ref/libgridsub_ref.so (the role cpp_subsampling/wrapper.cpp:58-285 plays for CPython).
#include <cstring>
#include <vector>
#include "grid_subsampling/grid_subsampling.h"

extern "C" int ref_grid_subsampling(const float* pts, const float* feats, const int* labels, int N, int fdim, int ldim,
                                    float dl, float* out_pts, float* out_feats, int* out_labels)
{
    std::vector<PointXYZ> in(N), out;
    for (int i = 0; i < N; ++i) in[i] = PointXYZ(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    std::vector<float> f_in, f_out;
    std::vector<int> l_in, l_out;
    if (fdim > 0) f_in.assign(feats, feats + (size_t)N * fdim);
    if (ldim > 0) l_in.assign(labels, labels + (size_t)N * ldim);
    grid_subsampling(in, out, f_in, f_out, l_in, l_out, dl, 0);
    const int M = (int)out.size();
    for (int i = 0; i < M; ++i) { out_pts[3 * i] = out[i].x; out_pts[3 * i + 1] = out[i].y; out_pts[3 * i + 2] = out[i].z; }
    if (fdim > 0) std::memcpy(out_feats, f_out.data(), sizeof(float) * (size_t)M * fdim);
    if (ldim > 0) std::memcpy(out_labels, l_out.data(), sizeof(int) * (size_t)M * ldim);
    return M;
}
