/*
 * ri_b200.h — C ABI of libri_b200.so: the B200 (sm_100a) implementation of the data-parallel front end of the
 * rotation-invariant PVCNN feature extractor.
 *
 * This is the drop-in boundary.  Each entry point replaces one function of the reference's pybind11 module
 * `_multi_shape_pvcnn_backend` (/root/reference/PVCNN/modules/functional/src/bindings.cpp:13-56) or, for the
 * matcher and the edge gather, the Python block cited beside it.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference adds in PVCNN/modules/functional/backend.py.
 *
 * Conventions (all entry points):
 *   - plain device pointers + sizes; tensors are contiguous, channel-major [B, C, N] (points innermost), fp32 / int32;
 *   - the CALLER allocates every output and the workspace; nothing is allocated or freed inside;
 *   - the caller has made the right device current and passes its stream (`cudaStream_t` as void*); every kernel is
 *     enqueued on that stream and the call returns without synchronising;
 *   - return value: 0 on success, a positive cudaError_t if a launch failed, or a negative RI_ERR_* code for a bad
 *     argument.  Never calls exit() (the reference's CUDA_CHECK_ERRORS does, cuda_utils.cuh:28-37);
 *   - re-entrant and thread-safe: the only process-wide state is read-only after its first use — the environment knobs of
 *     DESIGN.md §11 (read once), the SM count per device, and the per-(kernel, device) record that a kernel's function
 *     attributes have been set (csrc/ri_common.cuh).  Several host threads may call into the library on several devices;
 *   - calls that share a WORKSPACE must be ordered on one stream (the workspace holds their intermediate tables).
 */
#ifndef RI_B200_H
#define RI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RI_OK 0
#define RI_ERR_BAD_ARG (-1)
#define RI_ERR_WORKSPACE (-2)
#define RI_ERR_UNSUPPORTED (-3)

/* ABI version of this header/library pair. */
int ri_abi_version(void);

/* Debug aid: a one-thread kernel that stores the device's nanosecond timer (%globaltimer) into *slot. */
int ri_debug_stamp(unsigned long long* slot, void* stream);

/* Debug aid for tests and tools: the environment knobs of DESIGN.md §11 are read once per process; this sets one of them
 * afterwards, by its variable name (e.g. "RI_DEVOX_STREAM", -1 = unset).  Not to be called while launches are in flight on
 * other threads.  RI_ERR_BAD_ARG for a name that cannot be changed at run time. */
int ri_debug_set_knob(const char* name, int value);

/* ---- k-nearest neighbours -------------------------------------------------------------------------------
 * One direction of knn_forward_cuda (knn/knn.cpp:6-25 -> KnnKernel knn/knn.cu:5-49): for each of the n points of
 * xyz1 [B,c,n] its k nearest (squared L2) among the m points of xyz2 [B,c,m].
 * dist1 [B,k,n] ascending along k, idx1 [B,k,n] indices into xyz2; unfilled slots are (10000.0f, 0).
 * Bit-exact against the reference incl. its tie rule (lower reference index first). */
int ri_knn_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
               float* dist1, int* idx1, void* stream);

/* ri_knn_f32 through the one-thread-per-query kernel whatever the shape (ri_knn_f32 itself takes the warp-per-query
 * kernel of csrc/knn_warp.cu for c == 3, k <= 32, m <= 1024).  Same results bit for bit; exists for the tests. */
int ri_knn_thread_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                      float* dist1, int* idx1, void* stream);

/* knn_forward_cuda itself: both directions (knn/knn.cu:81-87). dist2/idx2 are [B,k,m]. */
int ri_knn_bilateral_f32(const float* xyz1, const float* xyz2, int B, int c, int n, int m, int k,
                         float* dist1, float* dist2, int* idx1, int* idx2, void* stream);

/* The same search for scan-sized clouds (~50k points, c == 3): uniform hash grid + ring expansion instead of the full
 * scan, results bit-identical to ri_knn_f32 (same distance expression, candidates ordered by (distance, index)).
 * k <= 32.  workspace >= ri_knn_grid_workspace_bytes(B, n, m) bytes, 16-byte aligned. */
size_t ri_knn_grid_workspace_bytes(int B, int n, int m);
int ri_knn_grid_f32(const float* xyz1, const float* xyz2, int B, int n, int m, int k,
                    float* dist1, int* idx1, void* workspace, size_t workspace_bytes, void* stream);

/* knn_backward_cuda (knn/knn.cpp:27-52 -> KnnGradKernel knn/knn.cu:52-78), both directions.
 * gradxyz1 [B,c,n] and gradxyz2 [B,c,m] are overwritten. */
int ri_knn_backward_f32(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                        const int* idx1, const int* idx2, int B, int c, int n, int m, int k,
                        float* gradxyz1, float* gradxyz2, void* stream);

/* ---- point-pair features --------------------------------------------------------------------------------
 * spherical_ppf_forward (spherical_ppf/ppf.cpp:17-36 -> spherical_ppf_kernel ppf.cu:19-92), backend argument
 * order: coords = the points, center = the centres (functional/ppf.py:22 swaps the user-facing order).
 * All inputs [B,3,L]; feat [B,4,L] = (angle(d,n_c), angle(d,n_p), angle(n_c,n_p), ||d||), d = center - coords. */
int ri_ppf_f32(const float* coords, const float* center, const float* normals, const float* center_normal,
               int B, int L, float* feat, void* stream);

/* Fused neighbour gather + PPF for k-NN neighbourhoods: xyz, normals [B,3,N], idx [B,k,N] -> out [B,4,k,N];
 * centre = point i, point = neighbour idx[b,s,i].  Equals ri_ppf_f32 on the gathered/expanded columns. */
int ri_ppf_gather_f32(const float* xyz, const float* normals, const int* idx, int B, int N, int k,
                      float* out, void* stream);

/* ---- coordinate prologue of the voxelization modules -----------------------------------------------------
 * Voxelization.forward (PVCNN/modules/voxelization.py:16-35) / Spherical_Voxelization.forward
 * (PVCNN/modules/spherical_vox.py:14-23) after the per-cloud mean: centre, scale, clamp/round or divide by the
 * largest radius, in one kernel.  points [B,pstride,N] (pstride 3 = xyz, 6 = xyz|normal), mean [B,3] (the caller's
 * `coords.mean(2)`: its summation order defines the bits).  shape: 0 cube normalize=False, 1 cube normalize=True,
 * 2 spherical.  Outputs: norm_coords [B,3,N] fp32 (always), vox_coords [B,3,N] i32 (shape 0/1), and optionally the
 * de-interleaved xyz / normals planes [B,3,N] (null to skip).  norm_mode selects the association of the 3-term
 * radius sum (1 = (x*x + y*y) + z*z without contraction, the one matching torch norm). */
int ri_vox_prologue_f32(const float* points, int pstride, const float* mean, int B, int N, int r,
                        int shape, float eps, int norm_mode,
                        float* xyz, float* normals, float* norm_coords, int* vox_coords, void* stream);

/* Fused self-query k-NN + PPF: ri_knn_f32(xyz, xyz) followed by ri_ppf_gather_f32, in one kernel and bit-identical to
 * that pair.  xyz / normals address [3,N] planes per cloud, cloud b at base + b * cloud_stride floats: 3N for two
 * contiguous [B,3,N] arrays, 6N (with normals = xyz + 3N) for the interleaved [B,6,N] input batch the models receive.
 * dist / idx [B,k,N] may both be null; ppf [B,4,k,N].  N <= 2048 and k <= 32, else RI_ERR_UNSUPPORTED. */
int ri_knn_ppf_f32(const float* xyz, const float* normals, long long cloud_stride, int B, int N, int k,
                   float* dist, int* idx, float* ppf, void* stream);

/* De-interleave points [B,6,N] (xyz | normal) into contiguous xyz [B,3,N] and normals [B,3,N] — the `.contiguous()`
 * copies of inputs[:, :3, :] / inputs[:, 3:, :] the reference's wrappers make (functional/knn.py:11-12,
 * functional/ppf.py:16-19) — and, optionally, the point-major packing packed [B,N,8] = (x,y,z,nx,ny,nz,0,0) (16-byte
 * aligned) that ri_ppf_gather_packed_f32 reads.  Any of the three outputs may be null; one launch. */
int ri_split_xyz_normals_f32(const float* points, int B, int N, float* xyz, float* normals, float* packed, void* stream);

/* ri_ppf_gather_f32 on the packed cloud: no shared memory, neighbours fetched as two 16-byte gathers.  Same values. */
int ri_ppf_gather_packed_f32(const float* packed, const int* idx, int B, int N, int k, float* out, void* stream);

/* ---- voxelization ---------------------------------------------------------------------------------------
 * spherical_avg_voxelize_forward (spherical_voxelization/spherical_vox.cpp:17-46) and avg_voxelize_forward
 * (voxelization/vox.cpp:17-43).  feat [B,C,N]; coords [B,3,N] (fp32 normalised Cartesian for the spherical
 * variant, int32 voxel coordinates for the cube variant) -> out [B,C,r^3], ind [B,N] (-1 = undefined point),
 * cnt [B,r^3].  All three outputs are fully overwritten (no pre-zeroing needed).
 * workspace >= ri_voxelize_workspace_bytes(B, C, N, r) bytes of device memory, 16-byte aligned (per-cloud cell tables
 * plus the compact table of cell means [B][C][<=N]). */
size_t ri_voxelize_workspace_bytes(int B, int C, int N, int r);
int ri_sph_voxelize_f32(const float* feat, const float* coords, int B, int C, int N, int r,
                        float* out, int* ind, int* cnt, void* workspace, size_t workspace_bytes, void* stream);
int ri_cube_voxelize_f32(const float* feat, const int* coords, int B, int C, int N, int r,
                         float* out, int* ind, int* cnt, void* workspace, size_t workspace_bytes, void* stream);

/* Fused voxelize + DGCNN edge features: same as above and additionally edge [B,2C,N] =
 * cat(feat - out[:, :, ind] (0 where ind == -1), feat), i.e. what ri_voxel_edge_gather_f32 would produce from `out`,
 * emitted in the same pass (the voxelizer already holds each point's cell mean; the dense grid is not re-read).
 * Replaces voxelize + PVCNN/modules/pvconv.py:68-90. */
int ri_sph_voxelize_edge_f32(const float* feat, const float* coords, int B, int C, int N, int r,
                             float* out, int* ind, int* cnt, float* edge,
                             void* workspace, size_t workspace_bytes, void* stream);
int ri_cube_voxelize_edge_f32(const float* feat, const int* coords, int B, int C, int N, int r,
                              float* out, int* ind, int* cnt, float* edge,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Phase-by-phase form of the voxelizers, for schedules that overlap or pipeline the phases (the one-shot calls above run
 * the same three kernels back to back):
 *   prepare  ind [B,N] and the cell tables of ALL B clouds (one launch over the batch);
 *   means    the compact table of cell means of the clouds [b0, b1) and, when edge [B,2C,N] is non-null, their DGCNN edge
 *            features cat(feat - mean of own cell, feat) (pvconv.py:68-90);
 *   fill     out [B,C,r^3] / cnt [B,r^3] of the clouds [b0, b1) from the tables and means in the workspace.
 * out, cnt, edge are the whole-batch arrays; same workspace as the one-shot calls.  Returns RI_ERR_UNSUPPORTED when the
 * shape is outside the tiled path (N > 4096, r^3 % 4 != 0, outputs not 16-byte aligned): use the one-shot calls then. */
int ri_sph_voxelize_prepare_f32(const float* coords, int B, int C, int N, int r, int* ind,
                                void* workspace, size_t workspace_bytes, void* stream);
int ri_cube_voxelize_prepare_f32(const int* coords, int B, int C, int N, int r, int* ind,
                                 void* workspace, size_t workspace_bytes, void* stream);
int ri_voxelize_means_f32(const float* feat, int B, int C, int N, int r, int b0, int b1, float* edge,
                          void* workspace, size_t workspace_bytes, void* stream);
int ri_voxelize_fill_f32(int B, int C, int N, int r, int b0, int b1, float* out, int* cnt,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The prefix of the voxel branch in one launch: ri_vox_prologue_f32 + ri_*_voxelize_prepare_f32 + ri_voxelize_means_f32
 * for all B clouds (bit-identical outputs); ri_voxelize_fill_f32 then completes the voxelization.  Arguments as
 * ri_vox_prologue_f32 plus feat [B,C,N]; outputs norm_coords [B,3,N], vox_coords [B,3,N] (cube shapes), ind [B,N],
 * edge [B,2C,N] (nullable) and the workspace tables.  N <= 1024 and the tiled-path conditions, else RI_ERR_UNSUPPORTED.
 * norm_mode flag bits: 0x100 = compute the per-cloud mean inside the kernel (torch's reduction order; written to `mean`);
 * 0x200 = `edge` is [B,C,N] and receives only the (feat - mean of the point's cell) half of the edge features — the
 * other half is the caller's own input. */
int ri_vox_front_f32(const float* points, int pstride, const float* mean, const float* feat,
                     int B, int C, int N, int r, int shape, float eps, int norm_mode,
                     float* norm_coords, int* vox_coords, int* ind, float* edge,
                     void* workspace, size_t workspace_bytes, void* stream);

/* avg_voxelize_backward == spherical_avg_voxelize_backward (vox.cpp:54-78, vox.cu:87-111):
 * grad_x [B,C,N] = grad_y[b,c,ind] / cnt (0 for undefined points); grad_x fully overwritten. */
int ri_voxelize_backward_f32(const float* grad_y, const int* ind, const int* cnt, int B, int C, int N, int s,
                             float* grad_x, void* stream);

/* ---- devoxelization -------------------------------------------------------------------------------------
 * trilinear_devoxelize_forward (interpolate/trilinear_devox.cpp:18-55): coords [B,3,N] in grid units [0,r-1],
 * feat [B,C,r^3] -> outs [B,C,N], inds [B,8,N], wgts [B,8,N] (all fully overwritten). */
int ri_trilinear_devox_f32(const float* coords, const float* feat, int B, int C, int N, int r,
                           float* outs, int* inds, float* wgts, void* stream);

/* spherical_trilinear_devoxelize_forward (interpolate/spherical_trilinear_devox.cpp:19-56), index quirks kept.
 * coords = normalised Cartesian coords, g_inds [B,N] = spherical cell of each point. Requires r >= 4. */
int ri_sph_trilinear_devox_f32(const float* coords, const float* feat, const int* g_inds, int B, int C, int N, int r,
                               float* outs, int* inds, float* wgts, void* stream);

/* trilinear_devoxelize_backward / spherical_trilinear_devoxelize_backward (trilinear_devox.cu:120-163,
 * spherical_trilinear_devox.cu:150-194): grad_x [B,C,s] fully overwritten; skip_undefined != 0 skips points whose
 * inds[b,0,i] == -1 (the spherical variant). */
int ri_devox_backward_f32(const float* grad_y, const int* inds, const float* wgts, int B, int C, int N, int s,
                          int skip_undefined, float* grad_x, void* stream);

/* ---- DGCNN voxel-neighbour edge features (PVCNN/modules/pvconv.py:68-90) --------------------------------
 * avg [B,C,s], feat [B,C,N], inds [B,N] -> out [B,2C,N] = cat(feat - avg[:, :, inds] (0 where inds == -1), feat). */
int ri_voxel_edge_gather_f32(const float* avg, const float* feat, const int* inds, int B, int C, int N, int s,
                             float* out, void* stream);

/* ---- ball query + grouping (the neighbourhood the shipped models use for local features) ----------
 * ball_query_forward (ball_query/ball_query.cpp:6-30 -> ball_query_kernel ball_query.cu:19-50): for each of the M centres
 * the first U points (in index order) with 1e-5 < d^2 < radius^2; the first hit fills the whole row, a centre without
 * neighbours keeps zeros.  centers [B,3,M], points [B,3,N] -> neighbors [B,M,U] int32, fully overwritten.  Bit-exact. */
int ri_ball_query_f32(const float* centers, const float* points, int B, int N, int M, float radius, int U,
                      int* neighbors, void* stream);

/* The 'ppf' local features of the shipped models from the neighbour indices, without materialising the grouped tensors:
 * PVCNN/models/pvcnn_classify.py:252-270 + PVCNN/modules/ball_query.py:16-35 (d = c - (p_nbr - c), |d|, three clamped acos
 * of torch-ordered dot products, fp32).  points_* [B,3,N], centers_* [B,3,M], neighbors [B,M,U] -> out [B,4,U,M]. */
int ri_local_ppf_f32(const float* points_coords, const float* points_normals, const float* centers_coords,
                     const float* centers_normals, const int* neighbors, int B, int N, int M, int U,
                     float* out, void* stream);

/* The whole local-feature branch of the shipped models (pvcnn_classify.py:61-67, 252-271) from the ball-query indices:
 * local point-pair features -> SharedMLP(4 -> 32 -> 64) in eval mode -> max over the neighbours, in one kernel (layer 2 on
 * tcgen05 as a 3xTF32 split product; nothing between the indices and the result leaves the SM).  w1 [32,4], b1 [32],
 * w2 [64,32], b2 [64]: the two 1x1 convolutions with their BatchNorm (running statistics) folded in.  neighbors [B,M,U];
 * out [B,64,M].  U == 128, C1 == 32, C2 == 64 (the shipped shape), else RI_ERR_UNSUPPORTED. */
int ri_local_ppf_mlp_max_f32(const float* points_coords, const float* points_normals, const float* centers_coords,
                             const float* centers_normals, const int* neighbors, int B, int N, int M, int U,
                             const float* w1, const float* b1, int C1, const float* w2, const float* b2, int C2,
                             float* out, void* stream);

/* grouping_forward / grouping_backward (grouping/grouping.cu:18-44, 58-84): out [B,C,M,U] = feat[b, c, idx[b, m, u]];
 * grad_x [B,C,N] += grad_y scattered through idx (float atomics, as the reference; grad_x is zeroed first). */
int ri_grouping_f32(const float* feat, const int* idx, int B, int C, int N, int M, int U, float* out, void* stream);
int ri_grouping_backward_f32(const float* grad_y, const int* idx, int B, int C, int N, int M, int U,
                             float* grad_x, void* stream);

/* ---- mutual-nearest-neighbour descriptor matching (datasets/deepgmr_mn40.py:232-244) ---------------------
 * find_correspondence_one_pair for P independent (source, target) pairs.
 * desc1, desc2: fp32 descriptors, channel-major [P,C,n1] / [P,C,n2] (what the feature extractor emits,
 * pvcnn_classify.py:345) when point_major == 0, or point-major [P,n1,C] / [P,n2,C] (the numpy layout the
 * reference's meter passes, deepgmr_mn40.py:121) when point_major != 0.
 *   diff[i,j] = |f1_i|^2 + |f2_j|^2 - 2 f1_i.f2_j           (3xTF32 split-precision tcgen05 contraction)
 *   corr12 [P,n1] = argmin_j diff, corr21 [P,n2] = argmin_i diff   (lowest index on ties, as np.argmin)
 *   dist12 [P,n1] = diff[i, corr12[i]] recomputed as an fp32 FMA chain; NULL = indices only (what the reference's
 *                   method returns): the re-evaluation pass over the descriptors is skipped
 *   idx1, idx2 [P,n1]: the mutual matches (corr21[corr12[i]] == i) in ascending i, count [P] of them, -1 beyond.
 * workspace >= ri_mutual_nn_workspace_bytes(P, C, n1, n2) bytes of device memory. */
size_t ri_mutual_nn_workspace_bytes(int P, int C, int n1, int n2);
int ri_mutual_nn_tf32x3(const float* desc1, const float* desc2, int P, int C, int n1, int n2, int point_major,
                        int* corr12, int* corr21, float* dist12, int* idx1, int* idx2, int* count,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- barycentre grid subsampling ---------------------------------------------------------------
 * grid_subsampling() (cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:4-106, called through
 * cpp_subsampling/wrapper.cpp:58-285 and utils/grid_subsampleing.py:3-21): one barycentre per occupied cell of edge dl.
 * points [N,3] row-major (the reference's numpy layout), features [N,fdim] or NULL, labels [N,ldim] or NULL ->
 * out_points [<=N,3], out_features [<=N,fdim], out_labels [<=N,ldim] (allocate for N cells), *out_count = number of cells
 * (DEVICE memory; no host synchronisation inside).  Cells come out in ascending cell index (the reference's order is an
 * unordered_map's); sums run in original point order, so barycentres and features are bit-identical to the reference;
 * label ties go to the smallest label.  workspace >= ri_grid_subsample_workspace_bytes(N). */
size_t ri_grid_subsample_workspace_bytes(int N);
int ri_grid_subsample_f32(const float* points, const float* features, const int* labels, int N, int fdim, int ldim,
                          float dl, float* out_points, float* out_features, int* out_labels, int* out_count,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- correspondences -> rigid pose -> registration metrics --------------------------------------
 * What follows the matcher in the reference's meter (datasets/deepgmr_mn40.py:104-139): the solve is delegated there to
 * Open3D RANSAC / FGR (utils/open3d_func.py:34-75) or TEASER++ (deepgmr_mn40.py:175-230) on the host, one pair at a time.
 * Here: RANSAC over the mutual matches with the configuration the reference hands to Open3D (3-point samples,
 * edge-length similarity 0.9, sample residual and validation distance = voxel_size, max_iter hypotheses, point-to-point
 * least squares without scale), all pairs in one launch, then Horn's closed-form refit on the inliers.
 * src [P,n1,3], tgt [P,n2,3] point-major fp32; idx1, idx2 [P,ld], count [P] as ri_mutual_nn_tf32x3 returns them;
 * hyps == 0: least squares over all matches (Kabsch).  T [P,4,4] row-major fp32 (src -> tgt), inliers [P];
 * best: P * 8 bytes of scratch (may be NULL when hyps == 0). */
int ri_pose_from_matches_f32(const float* src, const float* tgt, const int* idx1, const int* idx2, const int* count,
                             int P, int n1, int n2, int ld, int hyps, float inlier_dist, float edge_similarity,
                             int refine_iters, unsigned long long seed, float* T, int* inliers,
                             unsigned long long* best, void* stream);

/* RE_TE_one_pair (deepgmr_mn40.py:152-164) + the point RMSE of MeterModelNet40_registration.update (:121-126), fp64:
 * gt, est [P,4,4] row-major fp32, pts [P,n,3] -> out [P,3] double = (rotation error in degrees, translation error, rmse). */
int ri_registration_metrics_f32(const float* gt, const float* est, const float* pts, int P, int n, double* out,
                                void* stream);

/* ---- 'change_coords' rotation-invariant preprocessing -------------------------------------------
 * The Python double loop of PVCNN_classifier.forward, rot_invariant_preprocess == 'change_coords'
 * (PVCNN/models/pvcnn_classify.py:153-184), as one kernel: per cloud a frame from the farthest point and the farthest
 * point not (anti)parallel to it (|cos| < 0.9), Gram-Schmidt, and the coordinates expressed in that frame.
 * coords [B,cstride,N] (cstride 3 or 6, first three planes used), mean [B,3] = torch's coords.mean(2), norm_mode as in
 * ri_vox_prologue_f32  ->  out [B,3,N], bases [B,3,3] (rows x, y, z; may be NULL), ok [B] (0 where the reference asserts). */
int ri_lrf_change_coords_f32(const float* coords, int cstride, const float* mean, int B, int N, int norm_mode,
                             float* out, float* bases, int* ok, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RI_B200_H */
