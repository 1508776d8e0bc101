/* visma_b200.h — C ABI of libvisma_b200.so: the B200 (sm_100a) implementation of VISMA's
 * orientation-constrained ICP alignment path and its render_depth rasteriser.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/Eigen types.  Each entry point
 * names the reference interface it replaces (paths relative to the reference checkout; O3D =
 * thirdparty/Open3D).  INTEGRATION.md shows the C++ adapter a VISMA maintainer adds on top.
 *
 * Conventions
 *   - every function returns VB200_OK (0) or a negative vb200_status; nothing throws across the ABI;
 *   - all host buffers are caller-owned; device memory lives behind the opaque handles;
 *   - ICP 4x4 matrices are ROW-MAJOR double[16] (sidesteps the EIGEN_DEFAULT_TO_ROW_MAJOR ABI trap,
 *     CMakeLists.txt:12); renderer matrices are COLUMN-MAJOR float[16] exactly as the reference hands
 *     them to glUniformMatrix4fv (render/renderer.cpp:274-281,338);
 *   - a handle owns one CUDA stream; calls on one handle are serialised by the caller, different
 *     handles may be used from different threads;
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry returns
 *     VB200_ERR_NO_DEVICE / VB200_ERR_CUDA.
 */
#ifndef VISMA_B200_H
#define VISMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_VERSION 100 /* 0.1.0 */

typedef enum {
    VB200_OK = 0,
    VB200_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, radius > grid cell ...) */
    VB200_ERR_NO_DEVICE = -2, /* no CUDA device / device index out of range */
    VB200_ERR_CUDA = -3,      /* CUDA runtime error; text via vb200_last_error() */
    VB200_ERR_NOMEM = -4,
    VB200_ERR_NORMALS = -5,   /* point-to-plane requested without normals on both clouds
                                 (O3D/src/Core/Registration/Registration.cpp:152-157) */
    VB200_ERR_DISTANCE = -6   /* max_correspondence_distance <= 0 (Registration.cpp:148-151) */
} vb200_status;

/* estimator kinds */
#define VB200_EST_P2P 0           /* TransformationEstimationPointToPoint == cicp::...PointToPoint4DoF
                                     (O3D/src/Core/Registration/TransformationEstimation.cpp:47-59,
                                      src/constrained_ICP.cpp:25-37) */
#define VB200_EST_P2PLANE 1       /* TransformationEstimationPointToPlane (TransformationEstimation.cpp:75-103) */
#define VB200_EST_P2PLANE_GRAVITY 2 /* 4-DoF (yaw about the gravity axis + translation) point-to-plane:
                                     the constraint the class name include/constrained_ICP.h:14 promises */

typedef struct vb200_scene vb200_scene_t;  /* a target cloud resident on one GPU + its NN grid */
typedef struct vb200_batch vb200_batch_t;  /* source clouds + ICP problems resident on that GPU */

int vb200_version(void);
const char *vb200_strerror(int status);
const char *vb200_last_error(void); /* thread-local text of the last CUDA failure */
int vb200_device_count(void);       /* 0 when no driver/GPU is present; never fails */
/* Device memory the library has freed stays in CUDA's stream-ordered pool of that device (so that the next scene or
 * batch does not pay cudaMalloc / cudaFree); this hands it back to the driver.  Live scenes and batches are untouched. */
int vb200_release_cached_memory(int device);

/* ---- scene: replaces KDTreeFlann::SetGeometry (O3D/src/Core/Geometry/KDTreeFlann.cpp:70-87,191-208),
 * which the reference re-runs inside EVERY RegistrationICP call (Registration.cpp:160-161).
 * xyz / nrm: n x 3 doubles laid out like std::vector<Eigen::Vector3d> (PointCloud.h:86-87); nrm nullable.
 * max_radius: largest correspondence distance that will be queried (grid cell size); must be > 0. */
int vb200_scene_create(const double *xyz, const double *nrm, int64_t n, double max_radius, int device,
                       vb200_scene_t **out);
int vb200_scene_destroy(vb200_scene_t *scene);
int vb200_scene_size(const vb200_scene_t *scene, int64_t *n_points, int64_t *n_coarse_cells,
                     int64_t *n_fine_cells, double *cell_size);
/* the handle's cudaStream_t, for callers that time with CUDA events */
void *vb200_scene_stream(const vb200_scene_t *scene);
int vb200_scene_sync(const vb200_scene_t *scene);

/* ---- KNN: replaces KDTreeFlann::SearchHybrid(query, radius, max_nn = 1) called once per source point
 * from Registration.cpp:62-72 (KDTreeFlann.cpp:165-189).  Semantics reproduced exactly: d2 is the double
 * ((dx^2)+dy^2)+dz^2; a neighbour is accepted iff d2 < (double)(float)(radius*radius); exact ties break
 * to the lowest target index.  out_idx[i] = -1 and out_d2[i] = 0 where none.  q_xyz: host, Q x 3. */
int vb200_knn1(vb200_scene_t *scene, const double *q_xyz, int64_t Q, double radius, int32_t *out_idx,
               double *out_d2);
/* same search on DEVICE-resident buffers, asynchronous on the scene's stream (the bandwidth-sweep entry):
 * d_q_xyz Q x 3 doubles, d_out_idx Q int32, d_out_d2 Q doubles. */
int vb200_knn1_device(vb200_scene_t *scene, const void *d_q_xyz, int64_t Q, double radius,
                      void *d_out_idx, void *d_out_d2);
/* The same operator by exhaustive search, without a scene handle: every (query, target) distance in the
 * reference's double arithmetic (flann::L2, O3D/3rdparty/flann/algorithms/dist.h:150-177), the target cloud
 * streamed through shared memory by the TMA engine.  Same outputs bit for bit as vb200_knn1 (same threshold
 * rule, ties to the lowest target index).  Meant for queries against a cloud nothing has indexed yet and as the
 * independent on-device check of the grid search.  Up to 15 queries it streams the cloud as it is (24 B per target
 * point once per 8 queries, every distance in double: HBM-bound for 1-2 queries, FP64-bound beyond); from 16 queries
 * it screens every pair in f32 on a packed copy of the cloud (3 FFMA + 1 compare per pair, a rigorous error band) and
 * evaluates in double only the pairs that can matter.  Limits: Q <= 524 280 per call; the device variant needs a
 * 16-byte aligned target pointer and is enqueued on `cuda_stream` (a cudaStream_t, NULL = default) — from 16
 * queries it waits once for that stream (the bounding box of cloud and queries is read back to size the band). */
int vb200_knn1_bruteforce(const double *tgt_xyz, int64_t n, const double *q_xyz, int64_t Q, double radius,
                          int device, int32_t *out_idx, double *out_d2);
int vb200_knn1_bruteforce_device(const void *d_tgt_xyz, int64_t n, const void *d_q_xyz, int64_t Q, double radius,
                                 int device, void *d_out_idx, void *d_out_d2, void *cuda_stream);

/* ---- ICP operator: replaces open3d::RegistrationICP (O3D/src/Core/Registration/Registration.h:102-107,
 * Registration.cpp:141-186) for a BATCH of B independent sources against one scene.
 * src_xyz / src_nrm: concatenated sources, src_offsets[B+1] delimits them (src_nrm nullable; only its
 * presence matters, as in the reference); init_T: B x 16; gravity_axis: used by VB200_EST_P2PLANE_GRAVITY
 * (nullable otherwise); rel_fitness / rel_rmse / max_iter: ICPConvergenceCriteria (Registration.h:46-68).
 * Outputs (each nullable): out_T B x 16, out_fitness[B], out_rmse[B], out_ncorr[B], out_iters[B],
 * out_corr: sum(M_b) x 2 int32 — problem b's pairs (source index local to b, target index) start at
 * 2*src_offsets[b], first out_ncorr[b] rows valid, ascending source index.
 * Error behaviour follows the reference: max_dist <= 0 -> VB200_ERR_DISTANCE and out_T = init_T;
 * point-to-plane without normals on both sides -> VB200_ERR_NORMALS and out_T = init_T. */
int vb200_icp_run(vb200_scene_t *scene, const double *src_xyz, const double *src_nrm,
                  const int64_t *src_offsets, int32_t B, const double *init_T, int estimator,
                  const double *gravity_axis, double max_dist, double rel_fitness, double rel_rmse,
                  int max_iter, double *out_T, double *out_fitness, double *out_rmse, int32_t *out_ncorr,
                  int32_t *out_iters, int32_t *out_corr);

/* The same operator split into resident pieces (what vb200_icp_run composes, and what bench.py's
 * device-resident leg times): upload clouds once, define P problems = (cloud id, init), run, fetch. */
int vb200_batch_create(vb200_scene_t *scene, const double *src_xyz, const double *src_nrm,
                       const int64_t *src_offsets, int32_t n_clouds, vb200_batch_t **out);
int vb200_batch_destroy(vb200_batch_t *batch);
int vb200_batch_set_problems(vb200_batch_t *batch, const int32_t *cloud_ids, const double *init_T,
                             int32_t P);
/* asynchronous on the scene's stream */
int vb200_batch_run(vb200_batch_t *batch, int estimator, const double *gravity_axis, double max_dist,
                    double rel_fitness, double rel_rmse, int max_iter);
/* synchronises, then copies results out (each pointer nullable) */
int vb200_batch_results(vb200_batch_t *batch, double *out_T, double *out_fitness, double *out_rmse,
                        int32_t *out_ncorr, int32_t *out_iters);
/* correspondences of problem p after the last run: out_corr M_p x 2, returns count in *out_k */
int vb200_batch_corr(vb200_batch_t *batch, int32_t p, int32_t *out_corr, int32_t *out_k);
/* number of kernel launches issued by this batch since creation (bench.py's gpu_launches) */
int64_t vb200_batch_launches(const vb200_batch_t *batch);
/* n_iter unconditional ICP iterations (correspondence pass + estimator update, Registration.cpp:172-178)
 * continuing from the problems' current transforms, with no convergence test: the unit bench.py times.
 * Asynchronous on the scene's stream. */
int vb200_batch_iterate(vb200_batch_t *batch, int estimator, const double *gravity_axis, double max_dist,
                        int n_iter);
/* measurement knobs of a batch (never needed for results):
 *   VB200_OPT_NN_CACHE 0      every point is searched in every pass (the cached-neighbour tests are skipped);
 *                             results are identical, only slower — the ablation bench.py reports;
 *   VB200_OPT_SPLIT_TIMING 1  vb200_batch_iterate(..., n_iter = 1) brackets the pass and the solve with CUDA
 *                             events for vb200_batch_last_kernel_ms (the events keep the kernels of the
 *                             iteration from overlapping: a diagnostic, off by default). */
#define VB200_OPT_NN_CACHE 1
#define VB200_OPT_SPLIT_TIMING 2
int vb200_batch_set_option(vb200_batch_t *batch, int option, int value);
/* device time (ms, CUDA events on the scene's stream) of the correspondence-pass kernel launches and of
 * the solve kernel launches issued by the most recent vb200_batch_iterate call made with
 * VB200_OPT_SPLIT_TIMING; synchronises. */
int vb200_batch_last_kernel_ms(vb200_batch_t *batch, float *pass_ms, float *solve_ms);

/* ---- one cloud sharded over several GPUs (ICPRefinement's single global transform, src/evaluation.cpp:
 * 244-274, at multi-GPU scale): an iteration split at the only point where ranks must exchange data.
 *   vb200_batch_pass    correspondence pass over THIS rank's shard + per-problem totals (P x 32 doubles:
 *                       the 27 normal-equation sums or the 16 moments, sum d2, count) left in device memory;
 *   (caller)            all-reduce (sum) of the totals buffer across ranks — e.g. ncclAllReduce, 256 B/problem;
 *   vb200_batch_solve   fitness / rmse / convergence test / estimator update from the combined totals.
 * Every rank sees identical totals and therefore applies the identical update: transforms stay consistent
 * without a broadcast.  npts_global[P] (nullable = local sizes): source points of each problem over ALL ranks
 * (the fitness denominator, Registration.cpp:91).  pass_index: 0 for the pass at the initial transform, then
 * 1, 2, ... ; iteration k's update is skipped once pass_index >= max_iter or the problem has converged.
 * vb200_batch_set_totals_buffer lets the caller own the totals buffer (e.g. a torch tensor handed to NCCL). */
int vb200_batch_pass(vb200_batch_t *batch, int estimator, double max_dist);
int vb200_batch_set_totals_buffer(vb200_batch_t *batch, void *d_totals);
void *vb200_batch_totals(vb200_batch_t *batch);
int vb200_batch_solve(vb200_batch_t *batch, int estimator, const double *gravity_axis, double max_dist,
                      double rel_fitness, double rel_rmse, int max_iter, int pass_index,
                      const int64_t *npts_global);

/* ---- estimator plug-in: replaces TransformationEstimation::ComputeTransformation(source, target,
 * corres) (O3D/src/Core/Registration/TransformationEstimation.h:51-66; VISMA's subclass
 * include/constrained_ICP.h:22-26), so the reference's own CPU ICP loop can drive the GPU estimator.
 * src m x 3, tgt / tgt_nrm n x 3 (tgt_nrm nullable for P2P), corr K x 2 (source idx, target idx). */
int vb200_estimate(const double *src_xyz, int64_t m, const double *tgt_xyz, const double *tgt_nrm,
                   int64_t n, const int32_t *corr, int64_t K, int estimator, const double *gravity_axis,
                   int device, double out_T[16]);

/* The same with every cloud already resident on the GPU (d_src_xyz m x 3, d_tgt_xyz / d_tgt_nrm n x 3 doubles,
 * d_corr K x 2 int32, all DEVICE pointers): the K rows are gathered by the kernel, nothing is marshalled on the
 * host.  Runs on `cuda_stream` (a cudaStream_t, NULL = default) and synchronises it to return out_T.  Indices in
 * d_corr are trusted (the host entry validates them). */
int vb200_estimate_device(const void *d_src_xyz, int64_t m, const void *d_tgt_xyz, const void *d_tgt_nrm,
                          int64_t n, const void *d_corr, int64_t K, int estimator, const double *gravity_axis,
                          int device, void *cuda_stream, double out_T[16]);

/* ---- replaces cicp::TransformationEstimationPointToPoint4DoF::ComputeRMSE (src/constrained_ICP.cpp:13-23) and the
 * identical TransformationEstimationPointToPoint::ComputeRMSE (O3D/src/Core/Registration/TransformationEstimation.cpp:
 * 35-45): sqrt(sum |s_i - t_j|^2 / K) over the correspondences; 0 when K = 0.  (The ICP loop itself never calls
 * it: inlier_rmse_ comes from the search's distances, Registration.cpp:68,93.  The point-to-plane override,
 * TransformationEstimation.cpp:61-73, keeps only its LAST residual — `err = r * r` — and is not reproduced.) */
int vb200_rmse(const double *src_xyz, int64_t m, const double *tgt_xyz, int64_t n, const int32_t *corr, int64_t K,
               int device, double *out_rmse);

/* ---- orientation-constrained driver: replaces feh::RegisterModelToScene (src/annotation.cpp:29-64):
 * `level` yaw initialisations Ry(2*pi*i/level), one ICP each (all batched on the GPU), keep the run with
 * strictly the most correspondences (first wins ties).  point_to_plane selects VB200_EST_P2PLANE, else
 * VB200_EST_P2P (what cicp::...PointToPoint4DoF computes).  Criteria = ICPConvergenceCriteria(). */
int vb200_register_model_to_scene(vb200_scene_t *scan, const double *model_xyz, const double *model_nrm,
                                  int64_t m, int level, double threshold, int point_to_plane,
                                  double out_T[16], int32_t *out_ncorr, int32_t *out_best_level);

/* ---- render operator: replaces feh::Renderer::{SetCamera x2, SetMesh, RenderDepth}
 * (render/renderer.h:43-79, render/renderer.cpp:232-351) for a batch of meshes.
 * V_concat: vertices (x,y,z float) of all meshes, v_off[n_mesh+1]; F_concat: triangles (3 int32, indices
 * local to the mesh), f_off[n_mesh+1]; model_T: n_mesh x 16 column-major; view_T: the camera pose passed to
 * SetCamera(pose) (initial-camera -> current-camera), column-major; intrinsics as SetCamera(zn,zf,fx,fy,cx,cy).
 * out_z24 (nullable): n_mesh x H x W uint32 24-bit z (2^24-1 = background);
 * out_depth (nullable): n_mesh x H x W float = z24/(2^24-1), what glReadPixels(GL_DEPTH_COMPONENT,
 * GL_FLOAT) returns (renderer.cpp:343); row 0 = image top. */
int vb200_render_depth_batch(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                             const int64_t *f_off, int32_t n_mesh, const float *model_T,
                             const float view_T[16], float zn, float zf, float fx, float fy, float cx,
                             float cy, int H, int W, int device, uint32_t *out_z24, float *out_depth);

/* Same, with two extras for pipelines that keep the maps on the GPU and for measurement: when
 * outputs_on_device != 0, out_z24 / out_depth are DEVICE pointers written in place (no D2H copy);
 * kernel_ms (nullable) receives the device time of the clear / rasterise / resolve launches (CUDA events). */
int vb200_render_depth_batch_ex(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                                const int64_t *f_off, int32_t n_mesh, const float *model_T,
                                const float view_T[16], float zn, float zf, float fx, float fy, float cx,
                                float cy, int H, int W, int device, uint32_t *out_z24, float *out_depth,
                                int outputs_on_device, float *kernel_ms);

/* ---- RenderEdge / RenderMask: replaces feh::Renderer::RenderEdge and RenderMask (render/renderer.cpp:353-433,
 * render/shaders/edge_detection.frag:38-76): the depth pass above followed by the edge-detection pass on the
 * z-buffer.  edge_z_near / edge_z_far are the edge SHADER's linearisation uniforms, which the reference fixes
 * at 0.05 / 2.0 when it builds the shader (renderer.cpp:95-96) independently of the camera.  out_edge /
 * out_mask: n_mesh x H x W uint8 (each nullable); edge = round(255 * soft-thresholded mean neighbour depth
 * difference), 0 on a 5-pixel border and on background; mask follows the reference's polarity: 255 on
 * BACKGROUND — RenderMask clears the colour buffer to 1.0 and reads GL_RED back (render/renderer.cpp:411-422) —
 * and 0 where the mesh covers the pixel (undefined in the reference, whose depth shader writes no colour;
 * defined here as 0 so that the map is the binary mask the header promises, render/renderer.h:93). */
int vb200_render_edge_mask_batch(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                                 const int64_t *f_off, int32_t n_mesh, const float *model_T,
                                 const float view_T[16], float zn, float zf, float fx, float fy, float cx,
                                 float cy, int H, int W, int device, float edge_z_near, float edge_z_far,
                                 uint8_t *out_edge, uint8_t *out_mask, int outputs_on_device);

/* ---- voxel down-sample: replaces open3d::VoxelDownSample (O3D/src/Core/Geometry/DownSample.cpp:179-220),
 * the step before every ICP (src/evaluation.cpp:258, src/annotation.cpp:112).  Returns the number of
 * voxels in *out_n; out_xyz / out_nrm sized for n points (nrm/out_nrm nullable).  Output ordered by voxel
 * index (the reference's order is unordered_map iteration order: compare as sets). */
int vb200_voxel_downsample(const double *xyz, const double *nrm, int64_t n, double voxel_size, int device,
                           double *out_xyz, double *out_nrm, int64_t *out_n);

/* ---- mesh surface sampling: replaces feh::SamplePointCloudFromMesh (include/geometry.h:29-64), which builds
 * every ICP source cloud (src/evaluation.cpp:250-256, src/annotation.cpp:126).  V: nV x 3 float, F: nF x 3
 * int32; writes exactly n_samples area-weighted surface points (double xyz) and, if out_nrm is non-null, the
 * unit face normal of each sample.  Reproducible from `seed` (counter-based Philox keyed by sample index).
 * The reference seeds from the wall clock, mis-indexes the chosen face by one and samples the triangle's
 * parallelogram; neither quirk is reproduced (see sample.cu), so parity is statistical. */
int vb200_sample_mesh(const float *V, int64_t nV, const int32_t *F, int64_t nF, int64_t n_samples,
                      uint64_t seed, int device, double *out_xyz, double *out_nrm);

#ifdef __cplusplus
}
#endif
#endif /* VISMA_B200_H */
