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

/// The gravity-constrained 4-DoF point-to-plane estimator (yaw about `gravity` + translation): the constraint
/// the class name at include/constrained_ICP.h:14 promises and the reference never implemented (its 4DoF class is
/// a copy of the point-to-point one).  An open3d::TransformationEstimation like any other: the reference's own
/// CPU loop can drive it (ComputeTransformation runs on the GPU), and visma_b200::RegistrationICP recognises it
/// and runs the whole loop on the GPU with VB200_EST_P2PLANE_GRAVITY.
class TransformationEstimationPointToPlaneGravity : public open3d::TransformationEstimationPointToPlane {
public:
    explicit TransformationEstimationPointToPlaneGravity(const Eigen::Vector3d &gravity = Eigen::Vector3d(0, 1, 0),
                                                         int device = 0)
        : device_(device) {
        gravity_[0] = gravity(0); gravity_[1] = gravity(1); gravity_[2] = gravity(2);  // +Y: src/annotation.cpp:43,84
    }
    Eigen::Matrix4d ComputeTransformation(const open3d::PointCloud &source, const open3d::PointCloud &target,
                                          const open3d::CorrespondenceSet &corres) const override {
        if (corres.empty() || !target.HasNormals()) return Eigen::Matrix4d::Identity();
        std::vector<int32_t> c(2 * corres.size());
        for (size_t i = 0; i < corres.size(); i++) { c[2 * i] = corres[i](0); c[2 * i + 1] = corres[i](1); }
        double T[16];
        int rc = vb200_estimate(Raw(source.points_), (int64_t)source.points_.size(), Raw(target.points_),
                                Raw(target.normals_), (int64_t)target.points_.size(), c.data(), (int64_t)corres.size(),
                                VB200_EST_P2PLANE_GRAVITY, gravity_, device_, T);
        return rc == VB200_OK ? FromRowMajor(T) : Eigen::Matrix4d::Identity();
    }
    const double *gravity() const { return gravity_; }

private:
    double gravity_[3];
    int device_;
};

inline int EstimatorKind(const open3d::TransformationEstimation &e) {
    if (dynamic_cast<const TransformationEstimationPointToPlaneGravity *>(&e)) return VB200_EST_P2PLANE_GRAVITY;
    return e.GetTransformationEstimationType() == open3d::TransformationEstimationType::PointToPlane
                   ? VB200_EST_P2PLANE
                   : VB200_EST_P2P;
}

inline const double *GravityAxis(const open3d::TransformationEstimation &e) {
    auto *g = dynamic_cast<const TransformationEstimationPointToPlaneGravity *>(&e);
    return g ? g->gravity() : nullptr;
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
    // The reference decides per call (Registration.cpp:152-157): a point-to-plane problem whose source has no
    // normals (an empty source never has, PointCloud.h:74-76) returns RegistrationResult(init) — that one only.
    // Those sources are left out of the launch; the others run with `has normals` set.
    const bool plane = EstimatorKind(estimation) != VB200_EST_P2P;
    std::vector<int> run;  // positions in `sources` that go to the GPU
    for (int b = 0; b < B; b++) {
        if (plane && !sources[b]->HasNormals())
            open3d::PrintError("Error: TransformationEstimationPointToPlane requires pre-computed normal vectors.\n");
        else
            run.push_back(b);
    }
    const int R = (int)run.size();
    if (R == 0) return out;
    std::vector<int64_t> off(R + 1, 0);
    for (int k = 0; k < R; k++) off[k + 1] = off[k] + (int64_t)sources[run[k]]->points_.size();
    std::vector<double> xyz(3 * (size_t)off[R] + 3), T0(16 * (size_t)R), T(16 * (size_t)R), fit(R), rmse(R);
    for (int k = 0; k < R; k++) {
        const open3d::PointCloud &s = *sources[run[k]];
        if (!s.points_.empty()) std::copy(Raw(s.points_), Raw(s.points_) + 3 * s.points_.size(), xyz.begin() + 3 * off[k]);
        ToRowMajor(inits[run[k]], T0.data() + 16 * k);
    }
    std::vector<int32_t> nc(R), it(R), corr(2 * (size_t)off[R] + 2);
    // only the PRESENCE of source normals matters below the ABI (the estimator reads the target's)
    int rc = vb200_icp_run(target.handle(), xyz.data(), plane ? xyz.data() : nullptr, off.data(), R, T0.data(),
                           EstimatorKind(estimation), GravityAxis(estimation), max_correspondence_distance,
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
    for (int k = 0; k < R; k++) {
        open3d::RegistrationResult &r = out[run[k]];
        r.transformation_ = FromRowMajor(T.data() + 16 * k);
        r.fitness_ = fit[k];
        r.inlier_rmse_ = rmse[k];
        r.correspondence_set_.resize(nc[k]);
        for (int j = 0; j < nc[k]; j++)
            r.correspondence_set_[j] = Eigen::Vector2i(corr[2 * (off[k] + j)], corr[2 * (off[k] + j) + 1]);
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
    /// src/constrained_ICP.cpp:13-23 on the GPU (vb200_rmse); the ICP loop itself never calls it
    double ComputeRMSE(const PointCloud &source, const PointCloud &target,
                       const CorrespondenceSet &corres) const override {
        if (corres.empty()) return 0.0;
        std::vector<int32_t> c(2 * corres.size());
        for (size_t i = 0; i < corres.size(); i++) { c[2 * i] = corres[i][0]; c[2 * i + 1] = corres[i][1]; }
        double rmse = 0.0;
        int rc = vb200_rmse(visma_b200::Raw(source.points_), (int64_t)source.points_.size(),
                            visma_b200::Raw(target.points_), (int64_t)target.points_.size(), c.data(),
                            (int64_t)corres.size(), device_, &rmse);
        return rc == VB200_OK ? rmse : TransformationEstimationPointToPoint4DoF::ComputeRMSE(source, target, corres);
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
