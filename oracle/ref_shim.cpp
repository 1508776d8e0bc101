// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" façade over the UNMODIFIED reference sources, compiled
// where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libvisma_ref.so (git-ignored).  It lets the Python tests and
// bench.py's cpu_baseline / --impl reference legs call the real Open3D 0.3.0
// RegistrationICP / KDTreeFlann / VoxelDownSample and VISMA's
// cicp::TransformationEstimationPointToPoint4DoF on raw buffers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
// load this library.  Nothing under visma_b200/ links or dlopens it.
//
// Reference entry points wrapped (paths relative to /root/reference):
//   open3d::RegistrationICP            thirdparty/Open3D/src/Core/Registration/Registration.cpp:141-186
//   open3d::EvaluateRegistration       thirdparty/Open3D/src/Core/Registration/Registration.cpp:127-139
//   open3d::KDTreeFlann::SearchHybrid  thirdparty/Open3D/src/Core/Geometry/KDTreeFlann.cpp:165-189
//   TransformationEstimation*::ComputeTransformation
//                                      thirdparty/Open3D/src/Core/Registration/TransformationEstimation.cpp:47-103
//   cicp::...PointToPoint4DoF          src/constrained_ICP.cpp:13-37 (+ the one-method subclass the
//                                      reference forgot, include/constrained_ICP.h:14-30)
//   open3d::VoxelDownSample            thirdparty/Open3D/src/Core/Geometry/DownSample.cpp:179-220
//   feh::RegisterModelToScene          src/annotation.cpp:29-64 (stale TU: its 24-yaw loop is
//                                      restated here around the real RegistrationICP)
//
// All 4x4 matrices cross this ABI as row-major double[16].

#include <cstdint>
#include <cstring>
#include <cmath>
#include <memory>
#include <vector>
#include <chrono>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Core/Geometry/PointCloud.h>
#include <Core/Geometry/KDTreeFlann.h>
#include <Core/Registration/Registration.h>
#include <Core/Registration/TransformationEstimation.h>
#include "constrained_ICP.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// The reference class is abstract against the vendored Open3D;
// this adds the one missing override without touching reference files.
class Cicp4DoF : public open3d::cicp::TransformationEstimationPointToPoint4DoF {
public:
    open3d::TransformationEstimationType GetTransformationEstimationType()
            const override {
        return open3d::TransformationEstimationType::PointToPoint;
    }
};

void fill_cloud(open3d::PointCloud &pc, const double *xyz, const double *nrm,
                int64_t n) {
    pc.points_.resize(n);
    if (n) std::memcpy(pc.points_.data(), xyz, sizeof(double) * 3 * n);
    if (nrm) {
        pc.normals_.resize(n);
        if (n) std::memcpy(pc.normals_.data(), nrm, sizeof(double) * 3 * n);
    }
}

Eigen::Matrix4d from_rowmajor(const double *m) {
    Eigen::Matrix4d T;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) T(r, c) = m[4 * r + c];
    return T;
}

void to_rowmajor(const Eigen::Matrix4d &T, double *m) {
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m[4 * r + c] = T(r, c);
}

std::unique_ptr<open3d::TransformationEstimation> make_estimator(int kind) {
    switch (kind) {
        case 0:
            return std::unique_ptr<open3d::TransformationEstimation>(
                    new open3d::TransformationEstimationPointToPoint(false));
        case 1:
            return std::unique_ptr<open3d::TransformationEstimation>(
                    new open3d::TransformationEstimationPointToPlane());
        case 2:
            return std::unique_ptr<open3d::TransformationEstimation>(
                    new Cicp4DoF());
        default:
            return nullptr;
    }
}

void export_result(const open3d::RegistrationResult &res, double *out_T,
                   double *out_fitness, double *out_rmse, int32_t *out_ncorr,
                   int32_t *out_corr) {
    if (out_T) to_rowmajor(res.transformation_, out_T);
    if (out_fitness) *out_fitness = res.fitness_;
    if (out_rmse) *out_rmse = res.inlier_rmse_;
    if (out_ncorr) *out_ncorr = (int32_t)res.correspondence_set_.size();
    if (out_corr) {
        for (size_t i = 0; i < res.correspondence_set_.size(); i++) {
            out_corr[2 * i + 0] = res.correspondence_set_[i][0];
            out_corr[2 * i + 1] = res.correspondence_set_[i][1];
        }
    }
}

struct RefTree {
    open3d::PointCloud cloud;
    open3d::KDTreeFlann tree;
};

}  // namespace

extern "C" {

int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// open3d::RegistrationICP verbatim.  estimator: 0=p2p 1=p2plane 2=cicp 4DoF.
int ref_registration_icp(const double *src, const double *src_nrm, int64_t m,
                         const double *tgt, const double *tgt_nrm, int64_t n,
                         double max_dist, const double *init_T, int estimator,
                         double rel_fitness, double rel_rmse, int max_iter,
                         double *out_T, double *out_fitness, double *out_rmse,
                         int32_t *out_ncorr, int32_t *out_corr) {
    auto est = make_estimator(estimator);
    if (!est) return -1;
    open3d::PointCloud source, target;
    fill_cloud(source, src, src_nrm, m);
    fill_cloud(target, tgt, tgt_nrm, n);
    open3d::RegistrationResult res = open3d::RegistrationICP(
            source, target, max_dist, from_rowmajor(init_T), *est,
            open3d::ICPConvergenceCriteria(rel_fitness, rel_rmse, max_iter));
    export_result(res, out_T, out_fitness, out_rmse, out_ncorr, out_corr);
    return 0;
}

// open3d::EvaluateRegistration verbatim.
int ref_evaluate_registration(const double *src, int64_t m, const double *tgt,
                              int64_t n, double max_dist, const double *T,
                              double *out_fitness, double *out_rmse,
                              int32_t *out_ncorr, int32_t *out_corr) {
    open3d::PointCloud source, target;
    fill_cloud(source, src, nullptr, m);
    fill_cloud(target, tgt, nullptr, n);
    open3d::RegistrationResult res = open3d::EvaluateRegistration(
            source, target, max_dist, from_rowmajor(T));
    export_result(res, nullptr, out_fitness, out_rmse, out_ncorr, out_corr);
    return 0;
}

// Persistent KDTreeFlann for the correspondence-pass probes.
void *ref_kdtree_create(const double *tgt, int64_t n) {
    RefTree *t = new RefTree();
    fill_cloud(t->cloud, tgt, nullptr, n);
    if (!t->tree.SetGeometry(t->cloud)) {
        delete t;
        return nullptr;
    }
    return t;
}

void ref_kdtree_destroy(void *h) { delete (RefTree *)h; }

// SearchHybrid(query, radius, max_nn=1) per query, OpenMP over queries the way
// Registration.cpp:53-85 does.  out_idx = -1 where nothing is within radius.
int ref_kdtree_search_hybrid1(void *h, const double *q, int64_t nq,
                              double radius, int32_t *out_idx,
                              double *out_d2) {
    RefTree *t = (RefTree *)h;
    if (!t) return -1;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < nq; i++) {
        std::vector<int> indices(1);
        std::vector<double> dists(1);
        Eigen::Vector3d p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        if (t->tree.SearchHybrid(p, radius, 1, indices, dists) > 0) {
            out_idx[i] = indices[0];
            out_d2[i] = dists[0];
        } else {
            out_idx[i] = -1;
            out_d2[i] = 0.0;
        }
    }
    return 0;
}

// SearchKNN(query, 1) per query (the 50x50 golden test path,
// UnitTest/Core/Geometry/PointCloud.cpp:1074-1111).
int ref_kdtree_search_knn1(void *h, const double *q, int64_t nq,
                           int32_t *out_idx, double *out_d2) {
    RefTree *t = (RefTree *)h;
    if (!t) return -1;
    for (int64_t i = 0; i < nq; i++) {
        std::vector<int> indices(1);
        std::vector<double> dists(1);
        Eigen::Vector3d p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        if (t->tree.SearchKNN(p, 1, indices, dists) > 0) {
            out_idx[i] = indices[0];
            out_d2[i] = dists[0];
        } else {
            out_idx[i] = -1;
            out_d2[i] = 0.0;
        }
    }
    return 0;
}

// estimation.ComputeTransformation(source, target, corres) verbatim.
int ref_estimate(const double *src, int64_t m, const double *tgt,
                 const double *tgt_nrm, int64_t n, const int32_t *corr,
                 int64_t k, int estimator, double *out_T) {
    auto est = make_estimator(estimator);
    if (!est) return -1;
    open3d::PointCloud source, target;
    fill_cloud(source, src, nullptr, m);
    fill_cloud(target, tgt, tgt_nrm, n);
    open3d::CorrespondenceSet cs(k);
    for (int64_t i = 0; i < k; i++)
        cs[i] = Eigen::Vector2i(corr[2 * i], corr[2 * i + 1]);
    to_rowmajor(est->ComputeTransformation(source, target, cs), out_T);
    return 0;
}

// The ICP loop of Registration.cpp:159-185 re-driven from real Open3D pieces
// (persistent KDTreeFlann + SearchHybrid + estimator + PointCloud::Transform)
// so the per-iteration (fitness, rmse, T) trace is observable.  The final
// row equals ref_registration_icp's output (checked in tests/test_oracle_ref.py).
// trace rows: [fitness, rmse, ncorr, T(16)] = 19 doubles, rows 0..iters
// (row 0 = result at init).  Returns number of estimator iterations run.
int ref_icp_trace(void *tree_h, const double *src, const double *src_nrm,
                  int64_t m, const double *tgt_nrm, double max_dist,
                  const double *init_T, int estimator, double rel_fitness,
                  double rel_rmse, int max_iter, double *trace) {
    RefTree *t = (RefTree *)tree_h;
    auto est = make_estimator(estimator);
    if (!t || !est) return -1;
    open3d::PointCloud target_view;  // estimator needs normals on the target
    target_view.points_ = t->cloud.points_;
    if (tgt_nrm) {
        target_view.normals_.resize(t->cloud.points_.size());
        std::memcpy(target_view.normals_.data(), tgt_nrm,
                    sizeof(double) * 3 * t->cloud.points_.size());
    }
    open3d::PointCloud pcd;
    fill_cloud(pcd, src, src_nrm, m);
    Eigen::Matrix4d T = from_rowmajor(init_T);
    if (!T.isIdentity()) pcd.Transform(T);

    auto pass = [&](open3d::CorrespondenceSet &cs, double &fit, double &rmse) {
        std::vector<int32_t> idx(m);
        std::vector<double> d2(m);
        ref_kdtree_search_hybrid1(t, (const double *)pcd.points_.data(), m,
                                  max_dist, idx.data(), d2.data());
        cs.clear();
        double e2 = 0.0;
        for (int64_t i = 0; i < m; i++)
            if (idx[i] >= 0) {
                cs.push_back(Eigen::Vector2i((int)i, idx[i]));
                e2 += d2[i];
            }
        if (cs.empty()) {
            fit = rmse = 0.0;
        } else {
            fit = (double)cs.size() / (double)m;
            rmse = std::sqrt(e2 / (double)cs.size());
        }
    };
    auto put = [&](int row, double fit, double rmse, size_t k) {
        double *r = trace + 19 * row;
        r[0] = fit;
        r[1] = rmse;
        r[2] = (double)k;
        to_rowmajor(T, r + 3);
    };
    open3d::CorrespondenceSet cs;
    double fit, rmse;
    pass(cs, fit, rmse);
    put(0, fit, rmse, cs.size());
    int it = 0;
    for (; it < max_iter;) {
        Eigen::Matrix4d update = est->ComputeTransformation(pcd, target_view, cs);
        T = update * T;
        pcd.Transform(update);
        double pf = fit, pr = rmse;
        pass(cs, fit, rmse);
        it++;
        put(it, fit, rmse, cs.size());
        if (std::abs(pf - fit) < rel_fitness && std::abs(pr - rmse) < rel_rmse)
            break;
    }
    return it;
}

// open3d::VoxelDownSample verbatim.  Output order is the unordered_map's;
// callers compare as sets.  Returns number of output points (<= n).
int64_t ref_voxel_downsample(const double *xyz, const double *nrm, int64_t n,
                             double voxel, double *out_xyz, double *out_nrm) {
    open3d::PointCloud in;
    fill_cloud(in, xyz, nrm, n);
    auto out = open3d::VoxelDownSample(in, voxel);
    int64_t k = (int64_t)out->points_.size();
    if (out_xyz && k)
        std::memcpy(out_xyz, out->points_.data(), sizeof(double) * 3 * k);
    if (out_nrm && nrm && k)
        std::memcpy(out_nrm, out->normals_.data(), sizeof(double) * 3 * k);
    return k;
}

// feh::RegisterModelToScene (src/annotation.cpp:29-64): `level` yaw
// initialisations about +Y, keep the run with strictly more correspondences.
int ref_register_model_to_scene(const double *model, const double *model_nrm,
                                int64_t m, const double *scan,
                                const double *scan_nrm, int64_t n, int level,
                                double threshold, int point_to_plane,
                                double *out_T, int32_t *out_ncorr,
                                int32_t *out_best_level) {
    open3d::PointCloud source, target;
    fill_cloud(source, model, model_nrm, m);
    fill_cloud(target, scan, scan_nrm, n);
    double interval = 2 * M_PI / level;
    open3d::RegistrationResult best;
    int best_i = -1;
    for (int i = 0; i < level; ++i) {
        Eigen::Matrix4d init = Eigen::Matrix4d::Identity();
        init.block<3, 3>(0, 0) =
                Eigen::AngleAxis<double>(interval * i, Eigen::Vector3d::UnitY())
                        .toRotationMatrix();
        open3d::RegistrationResult r;
        if (point_to_plane) {
            r = open3d::RegistrationICP(
                    source, target, threshold, init,
                    open3d::TransformationEstimationPointToPlane(),
                    open3d::ICPConvergenceCriteria());
        } else {
            r = open3d::RegistrationICP(source, target, threshold, init,
                                        Cicp4DoF(),
                                        open3d::ICPConvergenceCriteria());
        }
        if (r.correspondence_set_.size() > best.correspondence_set_.size()) {
            best = r;
            best_i = i;
        }
    }
    to_rowmajor(best.transformation_, out_T);
    if (out_ncorr) *out_ncorr = (int32_t)best.correspondence_set_.size();
    if (out_best_level) *out_best_level = best_i;
    return 0;
}

// PointCloud::Transform (Geometry/PointCloud.cpp:75-87) on raw buffers.
int ref_transform(double *xyz, double *nrm, int64_t n, const double *T) {
    open3d::PointCloud pc;
    fill_cloud(pc, xyz, nrm, n);
    pc.Transform(from_rowmajor(T));
    if (n) std::memcpy(xyz, pc.points_.data(), sizeof(double) * 3 * n);
    if (nrm && n) std::memcpy(nrm, pc.normals_.data(), sizeof(double) * 3 * n);
    return 0;
}

}  // extern "C"
