// registration_b200.h — the C++ host side a VISMA build includes to run its ICP path on a B200.
//
// Header-only adapters with the reference's own signatures and Open3D types on top of the C ABI
// (include/visma_b200.h).  Everything crossing the ABI is copied element-wise through (row, col) accessors,
// so it works whether the including TU is compiled with -DEIGEN_DEFAULT_TO_ROW_MAJOR (VISMA,
// CMakeLists.txt:12) or without (Open3D) — the mismatch that breaks the reference's own link.
//
//   visma_b200::RegistrationICP(...)        same signature and error behaviour as open3d::RegistrationICP
//                                           (thirdparty/Open3D/src/Core/Registration/Registration.h:102-107)
//   visma_b200::RegistrationICPBatch(...)   B sources against one resident target in one launch sequence
//   visma_b200::RegisterModelToScene(...)   feh::RegisterModelToScene (src/annotation.cpp:29-64), 24 yaw inits batched
//   visma_b200::VoxelDownSample(...)        open3d::VoxelDownSample (Core/Geometry/DownSample.cpp:179-220)
//   open3d::cicp::TransformationEstimationPointToPoint4DoFB200
//                                           the reference's estimator class made instantiable (it forgets the
//                                           pure-virtual GetTransformationEstimationType, include/constrained_ICP.h:14-30)
//                                           with ComputeTransformation executed on the GPU
#pragma once

#include <cmath>
#include <memory>
#include <vector>

#include <Eigen/Core>
#include <Core/Geometry/PointCloud.h>
#include <Core/Registration/Registration.h>
#include <Core/Registration/TransformationEstimation.h>
#include <Core/Utility/Console.h>

#include "visma_b200.h"

namespace visma_b200 {

inline void ToRowMajor(const Eigen::Matrix4d &T, double *m) {
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m[4 * r + c] = T(r, c);
}

inline Eigen::Matrix4d FromRowMajor(const double *m) {
    Eigen::Matrix4d T;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) T(r, c) = m[4 * r + c];
    return T;
}

// std::vector<Eigen::Vector3d> is 24 bytes per point with no padding (PointCloud.h:86-88): usable as-is
inline const double *Raw(const std::vector<Eigen::Vector3d> &v) {
    return v.empty() ? nullptr : reinterpret_cast<const double *>(v.data());
}

inline int EstimatorKind(const open3d::TransformationEstimation &e) {
    return e.GetTransformationEstimationType() == open3d::TransformationEstimationType::PointToPlane
                   ? VB200_EST_P2PLANE
                   : VB200_EST_P2P;
}

/// A target cloud resident on the GPU.  The reference rebuilds its KD-tree inside every RegistrationICP call
/// (Registration.cpp:160-161); keep one of these alive across calls instead.
class Scene {
public:
    Scene(const open3d::PointCloud &target, double max_radius, int device = 0) {
        status_ = vb200_scene_create(Raw(target.points_), target.HasNormals() ? Raw(target.normals_) : nullptr,
                                     (int64_t)target.points_.size(), max_radius, device, &h_);
    }
    ~Scene() { vb200_scene_destroy(h_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    bool ok() const { return status_ == VB200_OK; }
    int status() const { return status_; }
    vb200_scene_t *handle() const { return h_; }

private:
    vb200_scene_t *h_ = nullptr;
    int status_ = VB200_OK;
};

/// B independent ICP problems against one scene — the batch form of open3d::RegistrationICP.
inline std::vector<open3d::RegistrationResult> RegistrationICPBatch(
        const std::vector<const open3d::PointCloud *> &sources, const Scene &target,
        double max_correspondence_distance, const std::vector<Eigen::Matrix4d> &inits,
        const open3d::TransformationEstimation &estimation =
                open3d::TransformationEstimationPointToPoint(false),
        const open3d::ICPConvergenceCriteria &criteria = open3d::ICPConvergenceCriteria()) {
    const int B = (int)sources.size();
    std::vector<open3d::RegistrationResult> out;
    for (int b = 0; b < B; b++) out.emplace_back(inits[b]);  // RegistrationResult(init), Registration.cpp:150,156
    if (B == 0) return out;
    std::vector<int64_t> off(B + 1, 0);
    bool normals = true;
    for (int b = 0; b < B; b++) {
        off[b + 1] = off[b] + (int64_t)sources[b]->points_.size();
        normals = normals && sources[b]->HasNormals();
    }
    std::vector<double> xyz(3 * (size_t)off[B]), T0(16 * (size_t)B), T(16 * (size_t)B), fit(B), rmse(B);
    for (int b = 0; b < B; b++) {
        std::copy(Raw(sources[b]->points_), Raw(sources[b]->points_) + 3 * sources[b]->points_.size(),
                  xyz.begin() + 3 * off[b]);
        ToRowMajor(inits[b], T0.data() + 16 * b);
    }
    std::vector<int32_t> nc(B), it(B), corr(2 * (size_t)off[B] + 2);
    int rc = vb200_icp_run(target.handle(), xyz.data(), normals ? xyz.data() : nullptr, off.data(), B, T0.data(),
                           EstimatorKind(estimation), nullptr, max_correspondence_distance,
                           criteria.relative_fitness_, criteria.relative_rmse_, criteria.max_iteration_, T.data(),
                           fit.data(), rmse.data(), nc.data(), it.data(), corr.data());
    if (rc == VB200_ERR_DISTANCE) {
        open3d::PrintError("Error: Invalid max_correspondence_distance.\n");
        return out;
    }
    if (rc == VB200_ERR_NORMALS) {
        open3d::PrintError("Error: TransformationEstimationPointToPlane requires pre-computed normal vectors.\n");
        return out;
    }
    if (rc != VB200_OK) {
        open3d::PrintError("visma_b200: %s (%s)\n", vb200_strerror(rc), vb200_last_error());
        return out;
    }
    for (int b = 0; b < B; b++) {
        out[b].transformation_ = FromRowMajor(T.data() + 16 * b);
        out[b].fitness_ = fit[b];
        out[b].inlier_rmse_ = rmse[b];
        out[b].correspondence_set_.resize(nc[b]);
        for (int k = 0; k < nc[b]; k++)
            out[b].correspondence_set_[k] = Eigen::Vector2i(corr[2 * (off[b] + k)], corr[2 * (off[b] + k) + 1]);
    }
    return out;
}

/// Drop-in for open3d::RegistrationICP: same arguments, same result type, same error behaviour.
inline open3d::RegistrationResult RegistrationICP(
        const open3d::PointCloud &source, const open3d::PointCloud &target, double max_correspondence_distance,
        const Eigen::Matrix4d &init = Eigen::Matrix4d::Identity(),
        const open3d::TransformationEstimation &estimation = open3d::TransformationEstimationPointToPoint(false),
        const open3d::ICPConvergenceCriteria &criteria = open3d::ICPConvergenceCriteria(), int device = 0) {
    if (max_correspondence_distance <= 0.0) {
        open3d::PrintError("Error: Invalid max_correspondence_distance.\n");
        return open3d::RegistrationResult(init);
    }
    Scene scene(target, max_correspondence_distance, device);
    if (!scene.ok()) {
        open3d::PrintError("visma_b200: %s (%s)\n", vb200_strerror(scene.status()), vb200_last_error());
        return open3d::RegistrationResult(init);
    }
    return RegistrationICPBatch({&source}, scene, max_correspondence_distance, {init}, estimation, criteria)[0];
}

/// feh::RegisterModelToScene(model, scene, options) (src/annotation.cpp:29-64).
inline Eigen::Matrix4d RegisterModelToScene(const open3d::PointCloud &model, const open3d::PointCloud &scan,
                                            int rotation_level, double distance_threshold, bool point_to_plane,
                                            int device = 0) {
    Eigen::Matrix4d I = Eigen::Matrix4d::Identity();
    Scene scene(scan, distance_threshold, device);
    if (!scene.ok()) return I;
    double T[16];
    int32_t nc = 0, best = -1;
    int rc = vb200_register_model_to_scene(scene.handle(), Raw(model.points_),
                                           model.HasNormals() ? Raw(model.normals_) : nullptr,
                                           (int64_t)model.points_.size(), rotation_level, distance_threshold,
                                           point_to_plane ? 1 : 0, T, &nc, &best);
    if (rc != VB200_OK && rc != VB200_ERR_NORMALS && rc != VB200_ERR_DISTANCE) return I;
    return FromRowMajor(T);
}

/// open3d::VoxelDownSample (output ordered by voxel index instead of unordered_map order).
inline std::shared_ptr<open3d::PointCloud> VoxelDownSample(const open3d::PointCloud &input, double voxel_size,
                                                          int device = 0) {
    auto output = std::make_shared<open3d::PointCloud>();
    if (voxel_size <= 0.0 || input.points_.empty()) return output;
    const bool nrm = input.HasNormals();
    output->points_.resize(input.points_.size());
    if (nrm) output->normals_.resize(input.points_.size());
    int64_t k = 0;
    int rc = vb200_voxel_downsample(Raw(input.points_), nrm ? Raw(input.normals_) : nullptr,
                                    (int64_t)input.points_.size(), voxel_size, device,
                                    reinterpret_cast<double *>(output->points_.data()),
                                    nrm ? reinterpret_cast<double *>(output->normals_.data()) : nullptr, &k);
    if (rc != VB200_OK) k = 0;
    output->points_.resize(k);
    if (nrm) output->normals_.resize(k);
    return output;
}

/// feh::SamplePointCloudFromMesh (include/geometry.h:29-64) on the GPU; V: n x 3 float row-major, F: m x 3 int.
/// Reproducible from `seed` (the reference seeds from the wall clock).
inline std::vector<Eigen::Vector3d> SamplePointCloudFromMesh(const float *V, int64_t nV, const int *F, int64_t nF,
                                                             int max_num_pts = 1000, uint64_t seed = 0,
                                                             int device = 0) {
    std::vector<Eigen::Vector3d> out((size_t)std::max(max_num_pts, 0));
    if (out.empty()) return out;
    int rc = vb200_sample_mesh(V, nV, F, nF, max_num_pts, seed, device, reinterpret_cast<double *>(out.data()),
                               nullptr);
    if (rc != VB200_OK) out.clear();
    return out;
}

}  // namespace visma_b200

#ifdef VISMA_B200_WITH_CICP
#include "constrained_ICP.h"
namespace open3d {
namespace cicp {

/// The reference's estimator, instantiable, with ComputeTransformation on the GPU: lets the reference's own
/// CPU RegistrationICP loop (Registration.cpp:172) drive the CUDA estimator unchanged.
class TransformationEstimationPointToPoint4DoFB200 : public TransformationEstimationPointToPoint4DoF {
public:
    explicit TransformationEstimationPointToPoint4DoFB200(int device = 0) : device_(device) {}
    TransformationEstimationType GetTransformationEstimationType() const override {
        return TransformationEstimationType::PointToPoint;
    }
    Eigen::Matrix4d ComputeTransformation(const PointCloud &source, const PointCloud &target,
                                          const CorrespondenceSet &corres) const override {
        if (corres.empty()) return Eigen::Matrix4d::Identity();
        std::vector<int32_t> c(2 * corres.size());
        for (size_t i = 0; i < corres.size(); i++) { c[2 * i] = corres[i][0]; c[2 * i + 1] = corres[i][1]; }
        double T[16];
        int rc = vb200_estimate(visma_b200::Raw(source.points_), (int64_t)source.points_.size(),
                                visma_b200::Raw(target.points_), nullptr, (int64_t)target.points_.size(), c.data(),
                                (int64_t)corres.size(), VB200_EST_P2P, nullptr, device_, T);
        return rc == VB200_OK ? visma_b200::FromRowMajor(T) : Eigen::Matrix4d::Identity();
    }

private:
    int device_;
};

}  // namespace cicp
}  // namespace open3d
#endif
