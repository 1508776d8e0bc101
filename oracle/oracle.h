/* oracle/oracle.h — CPU restatement of the reference's ICP / KNN / depth-render path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may build, link, load or call anything declared here.
 * The shipped library (visma_b200/csrc) never includes this header.
 *
 * Parity status:
 *   - ICP / KNN / estimators: PINNED against (a) the Open3D docs known-answer test on
 *     cloud_bin_0 -> cloud_bin_1 (docs/tutorial/Basic/icp_registration.rst:56-58,91-98,154-161),
 *     (b) the 50x50 1-NN golden vector of UnitTest/Core/Geometry/PointCloud.cpp:1074-1111 and
 *     (c) outputs of the unmodified reference compiled here (oracle/_ref, tests/test_oracle_vs_ref.py).
 *   - depth rasteriser: PARITY UNPINNED by the reference — it ships no golden depth map and its
 *     OpenGL renderer (glm + vendor rasteriser) cannot run in this container.  The rules restated
 *     here are GL 3.3's (pixel-centre sampling, window-space-linear z, unorm24 GL_LESS) with a
 *     canonical sub-pixel snapping (1/256 px, top-left fill) that GL leaves to the vendor.
 *
 * All 4x4 ICP matrices are row-major double[16].  Renderer matrices are column-major
 * float[16] exactly as the reference passes them to glUniformMatrix4fv.
 * Paths below are relative to /root/reference; O3D = thirdparty/Open3D.
 */
#ifndef VISMA_ORACLE_H
#define VISMA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VO_P2P = 0, VO_P2PLANE = 1, VO_P2PLANE_GRAVITY = 2, VO_P2P_CICP = 3 };

/* ---- nearest-neighbour index over a target cloud --------------------------------------------- */
typedef struct vo_index vo_index;

/* Uniform-grid exact NN index (replaces FLANN's KD-tree; same answers, see vo_knn1). */
vo_index *vo_index_create(const double *tgt_xyz, int64_t n, double cell);
void vo_index_destroy(vo_index *ix);

/* KDTreeFlann::SearchHybrid(q, radius, max_nn=1) semantics (O3D/src/Core/Geometry/KDTreeFlann.cpp:165-189,
 * FLANN result_set.h:529-637, dist.h:150-177): j* = argmin_j ((dx^2)+dy^2)+dz^2 in double; accepted iff
 * d2 < (double)(float)(radius*radius).  Exact-distance ties break to the LOWEST target index (FLANN:
 * first visited in tree order — traversal dependent; measure-zero on real data).  out_idx = -1, out_d2 = 0
 * where nothing is within the radius.  `radius` must not exceed the index's cell size. */
int vo_knn1(const vo_index *ix, const double *q_xyz, int64_t nq, double radius,
            int32_t *out_idx, double *out_d2);

/* KDTreeFlann::SearchKNN(q, 1) semantics (no radius), brute force; small inputs only. */
int vo_knn1_brute(const double *tgt_xyz, int64_t n, const double *q_xyz, int64_t nq,
                  int32_t *out_idx, double *out_d2);

/* ---- estimators ----------------------------------------------------------------------------------- */
/* TransformationEstimationPointToPoint::ComputeTransformation == cicp::...4DoF::ComputeTransformation
 * (O3D/src/Core/Registration/TransformationEstimation.cpp:47-59, src/constrained_ICP.cpp:25-37):
 * Eigen::umeyama(src, dst, with_scaling) (O3D/3rdparty/Eigen/Eigen/src/Geometry/Umeyama.h:93-162). */
int vo_estimate_p2p(const double *src_xyz, const double *tgt_xyz, const int32_t *corr, int64_t k,
                    int with_scaling, double out_T[16]);

/* TransformationEstimationPointToPlane::ComputeTransformation
 * (TransformationEstimation.cpp:75-103; Utility/Eigen.cpp:35-68,88-106,137-182). */
int vo_estimate_p2plane(const double *src_xyz, const double *tgt_xyz, const double *tgt_nrm,
                        const int32_t *corr, int64_t k, double out_T[16]);

/* Gravity-constrained 4-DoF point-to-plane step (north-star extension, NOT in the reference;
 * SURVEY App. A): x = [theta, t]; J4 = [(vs x nt).g ; nt]; update.R = AngleAxis(theta, g). */
int vo_estimate_p2plane_gravity(const double *src_xyz, const double *tgt_xyz, const double *tgt_nrm,
                                const int32_t *corr, int64_t k, const double g[3], double out_T[16]);

/* ComputeRMSE of the p2p estimators (TransformationEstimation.cpp:35-45, src/constrained_ICP.cpp:13-23). */
double vo_rmse_p2p(const double *src_xyz, const double *tgt_xyz, const int32_t *corr, int64_t k);

/* SolveLinearSystem + TransformVector6dToMatrix4d exposed for unit tests (Utility/Eigen.cpp:35-68). */
int vo_solve6(const double JTJ[36], const double JTr[6], double out_x[6]); /* 1 = solved, 0 = det guard */
void vo_vec6_to_T(const double x[6], double out_T[16]);

/* ---- ICP loop ------------------------------------------------------------------------------------- */
/* open3d::RegistrationICP (O3D/src/Core/Registration/Registration.cpp:141-186) with
 * GetRegistrationResultAndCorrespondences (:41-96).  trace (nullable): (max_iter+1) rows of
 * [fitness, rmse, ncorr, T(16)]; row 0 = result at init.  out_corr (nullable): m x 2 int32, first
 * *out_ncorr rows valid, ascending source index.  Returns 0, or -1 on the reference's error paths
 * (max_dist <= 0, or p2plane without normals on both clouds: result = init, fitness = rmse = 0). */
int vo_registration_icp(const vo_index *ix, const double *tgt_xyz, const double *tgt_nrm, int64_t n,
                        const double *src_xyz, const double *src_nrm, int64_t m,
                        double max_dist, const double init_T[16], int estimator, const double gravity[3],
                        double rel_fitness, double rel_rmse, int max_iter,
                        double out_T[16], double *out_fitness, double *out_rmse,
                        int32_t *out_ncorr, int32_t *out_iters, int32_t *out_corr, double *trace);

/* feh::RegisterModelToScene (src/annotation.cpp:29-64): `level` yaw inits about +Y, keep the run with
 * strictly more correspondences (first wins ties). */
int vo_register_model_to_scene(const double *scan_xyz, const double *scan_nrm, int64_t n,
                               const double *model_xyz, const double *model_nrm, int64_t m,
                               int level, double threshold, int point_to_plane,
                               double out_T[16], int32_t *out_ncorr, int32_t *out_best_level);

/* ---- depth rasteriser (render/renderer.cpp:232-351, render/shaders/basic_mvp.vert:10) ----------- */
/* Renderer::SetCamera(zn,zf,fx,fy,cx,cy) projection (renderer.cpp:259-268), column-major float[16]. */
void vo_projection(float zn, float zf, float fx, float fy, float cx, float cy, int H, int W,
                   float out_P[16]);
/* Renderer::SetCamera(pose): view = diag(1,-1,-1,1) * pose (renderer.cpp:284-300). */
void vo_view(const float pose[16], float out_V[16]);
/* Renderer::RenderDepth for one mesh: out_z24 (nullable) H*W uint32 in [0, 2^24-1] (2^24-1 = background),
 * out_depth (nullable) H*W float = q/(2^24-1), row 0 = image top. */
int vo_render_depth(const float *V, int64_t nV, const int32_t *F, int64_t nF, const float model[16],
                    const float view[16], const float proj[16], int H, int W,
                    uint32_t *out_z24, float *out_depth);
/* Renderer::RenderEdge's screen pass on a z-buffer (render/shaders/edge_detection.frag:38-76; zn/zf are the
 * SHADER's uniforms, 0.05 / 2.0 in the reference, renderer.cpp:95-96) and RenderMask (z != 1 -> 255). */
int vo_render_edge(const uint32_t *z24, int H, int W, float zn, float zf, uint8_t *out_edge);
int vo_render_mask(const uint32_t *z24, int H, int W, uint8_t *out_mask);
/* LinearizeDepth (render/renderer.h:32-36) */
float vo_linearize_depth(float zb, float zn, float zf);

/* ---- voxel down-sample (O3D/src/Core/Geometry/DownSample.cpp:179-220) ---------------------------- */
/* Output in order of each voxel's first point in the input (the reference's order is unordered_map
 * iteration order: compare as sets).  Returns number of voxels, or -1 on the reference's error paths. */
int64_t vo_voxel_downsample(const double *xyz, const double *nrm, int64_t n, double voxel,
                            double *out_xyz, double *out_nrm);

#ifdef __cplusplus
}
#endif
#endif
