// tests/cpp/sharded_check.cpp — visma_b200::RegistrationICPSharded (visma_b200/host/sharded_b200.h) on the GPUs of
// this box: one host thread per GPU (ncclCommInitAll), each with its own scene replica, against the single-GPU
// RegistrationICPBatch of all the objects, and visma_b200::RegistrationICPGlobalSharded (one cloud cut into slices, a
// 256-byte ncclAllReduce per iteration) against the single-GPU alignment of the whole cloud.  TEST INFRASTRUCTURE; built by `make -C oracle sharded` where
// /root/reference exists (Open3D headers), run by tests/test_gpu_dropin.py.
//
//   usage: sharded_check <input.bin> [world]     world defaults to min(device count, 2)
//   input: int64 n_tgt, B, m; tgt xyz, tgt nrm; B x (src xyz, src nrm) of m points; B x init (16 doubles, row-major)
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "sharded_b200.h"

static double max_abs_diff(const Eigen::Matrix4d &a, const Eigen::Matrix4d &b) {
    double m = 0;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m = std::max(m, std::abs(a(r, c) - b(r, c)));
    return m;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t n[3];
    if (fread(n, sizeof(int64_t), 3, f) != 3) return 2;
    const int B = (int)n[1];
    open3d::PointCloud target;
    target.points_.resize(n[0]); target.normals_.resize(n[0]);
    bool ok = fread(target.points_.data(), 24, n[0], f) == (size_t)n[0] && fread(target.normals_.data(), 24, n[0], f) == (size_t)n[0];
    std::vector<open3d::PointCloud> src(B);
    for (int b = 0; b < B && ok; b++) {
        src[b].points_.resize(n[2]); src[b].normals_.resize(n[2]);
        ok = fread(src[b].points_.data(), 24, n[2], f) == (size_t)n[2] && fread(src[b].normals_.data(), 24, n[2], f) == (size_t)n[2];
    }
    std::vector<Eigen::Matrix4d> inits(B);
    for (int b = 0; b < B && ok; b++) {
        double m[16];
        ok = fread(m, 8, 16, f) == 16;
        inits[b] = visma_b200::FromRowMajor(m);
    }
    fclose(f);
    if (!ok) return 2;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 1) return 3;
    const int world = argc > 2 ? atoi(argv[2]) : std::min(ndev, 2);
    if (world > ndev) return 4;  // NCCL refuses two ranks on one device
    const double max_d = 0.075;
    open3d::TransformationEstimationPointToPlane p2l;
    std::vector<const open3d::PointCloud *> ptrs;
    for (auto &s : src) ptrs.push_back(&s);

    // single GPU, all objects in one batch
    std::vector<open3d::RegistrationResult> single;
    {
        visma_b200::Scene scene(target, max_d, 0);
        single = visma_b200::RegistrationICPBatch(ptrs, scene, max_d, inits, p2l);
    }
    // `world` ranks, one host thread and one scene replica per GPU
    std::vector<int> devs(world);
    for (int r = 0; r < world; r++) devs[r] = r;
    std::vector<ncclComm_t> comms(world);
    if (ncclCommInitAll(comms.data(), world, devs.data()) != ncclSuccess) return 5;
    std::vector<std::vector<open3d::RegistrationResult>> per_rank(world);
    std::vector<std::vector<int>> ncorr(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r]() {
            cudaSetDevice(devs[r]);
            visma_b200::Scene scene(target, max_d, devs[r]);
            per_rank[r] = visma_b200::RegistrationICPSharded(ptrs, scene, max_d, inits, p2l,
                                                             open3d::ICPConvergenceCriteria(), comms[r], r, world,
                                                             nullptr, &ncorr[r]);
        });
    for (auto &t : th) t.join();
    // the other split: ONE cloud (object 0) cut into `world` slices, one 256-byte all-reduce per iteration, against the
    // single-GPU alignment of the whole cloud
    open3d::RegistrationResult whole(inits[0]);
    {
        visma_b200::Scene scene(target, max_d, 0);
        whole = visma_b200::RegistrationICPBatch({&src[0]}, scene, max_d, {inits[0]}, p2l)[0];
    }
    std::vector<open3d::RegistrationResult> glob(world, open3d::RegistrationResult(inits[0]));
    th.clear();
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r]() {
            cudaSetDevice(devs[r]);
            visma_b200::Scene scene(target, max_d, devs[r]);
            const size_t m = src[0].points_.size(), a = r * m / world, e = (r + 1) * m / world;
            open3d::PointCloud slice;
            slice.points_.assign(src[0].points_.begin() + a, src[0].points_.begin() + e);
            slice.normals_.assign(src[0].normals_.begin() + a, src[0].normals_.begin() + e);
            glob[r] = visma_b200::RegistrationICPGlobalSharded(slice, scene, max_d, inits[0], p2l,
                                                               open3d::ICPConvergenceCriteria(), comms[r]);
        });
    for (auto &t : th) t.join();
    double g_dT = 0, g_dfit = 0, g_ranks = 0;
    for (int r = 0; r < world; r++) {
        g_dT = std::max(g_dT, max_abs_diff(glob[r].transformation_, whole.transformation_));
        g_dfit = std::max(g_dfit, std::abs(glob[r].fitness_ - whole.fitness_) + std::abs(glob[r].inlier_rmse_ - whole.inlier_rmse_));
        g_ranks = std::max(g_ranks, max_abs_diff(glob[r].transformation_, glob[0].transformation_));
    }
    for (auto &c : comms) ncclCommDestroy(c);
    double worst = 0, worst_fit = 0;
    int ncorr_bad = 0, own_sets = 0;
    for (int r = 0; r < world; r++)
        for (int b = 0; b < B; b++) {
            worst = std::max(worst, max_abs_diff(per_rank[r][b].transformation_, single[b].transformation_));
            worst_fit = std::max(worst_fit, std::abs(per_rank[r][b].fitness_ - single[b].fitness_) +
                                                    std::abs(per_rank[r][b].inlier_rmse_ - single[b].inlier_rmse_));
            ncorr_bad += ncorr[r][b] != (int)single[b].correspondence_set_.size();
            if (b % world == r) own_sets += per_rank[r][b].correspondence_set_ == single[b].correspondence_set_;
        }
    printf("\nJSON:{\"world\": %d, \"objects\": %d, \"max_dT\": %.3e, \"max_dfit\": %.3e, \"ncorr_mismatches\": %d, "
           "\"own_correspondence_sets_equal\": %d, \"fitness0\": %.6f, \"global_max_dT\": %.3e, \"global_max_dfit\": %.3e, "
           "\"global_rank_spread\": %.3e, \"global_fitness\": %.6f}\n",
           world, B, worst, worst_fit, ncorr_bad, own_sets, single[0].fitness_, g_dT, g_dfit, g_ranks, whole.fitness_);
    return 0;
}
