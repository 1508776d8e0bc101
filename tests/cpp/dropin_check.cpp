// tests/cpp/dropin_check.cpp — compiles the C++ host adapters (visma_b200/host/*.h) against the reference's
// real Open3D headers and runs them next to the reference's own CPU functions on the same PointCloud objects.
// TEST INFRASTRUCTURE: links oracle/_ref objects (the unmodified reference) AND libvisma_b200.so.
// Built by `make -C oracle dropin` where /root/reference exists; run by tests/test_gpu_dropin.py.
//
//   usage: dropin_check <input.bin>      prints one JSON object with the differences
//   input: int64 n_tgt, n_src; tgt xyz, tgt nrm, src xyz, src nrm (doubles); init (16 doubles, row-major)
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define VISMA_B200_WITH_CICP
#include <Eigen/LU>

#include "registration_b200.h"
#include "renderer_b200.h"

// Open3D's PrintError writes (coloured) text to stdout, so the JSON is collected and emitted last, alone on a line
static std::string g_json;
static void out(const char *fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_json += buf;
}

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
    int64_t n[2];
    if (fread(n, sizeof(int64_t), 2, f) != 2) return 2;
    open3d::PointCloud target, source;
    target.points_.resize(n[0]); target.normals_.resize(n[0]);
    source.points_.resize(n[1]); source.normals_.resize(n[1]);
    double init_rm[16];
    bool ok = fread(target.points_.data(), 24, n[0], f) == (size_t)n[0] &&
              fread(target.normals_.data(), 24, n[0], f) == (size_t)n[0] &&
              fread(source.points_.data(), 24, n[1], f) == (size_t)n[1] &&
              fread(source.normals_.data(), 24, n[1], f) == (size_t)n[1] && fread(init_rm, 8, 16, f) == 16;
    fclose(f);
    if (!ok) return 2;
    Eigen::Matrix4d init = visma_b200::FromRowMajor(init_rm);
    const double max_d = 0.075;

    out("{");
    // (1) the ICP operator: reference CPU vs the drop-in, both estimators
    open3d::TransformationEstimationPointToPoint p2p;
    open3d::TransformationEstimationPointToPlane p2l;
    const open3d::TransformationEstimation *ests[2] = {&p2p, &p2l};
    const char *names[2] = {"p2p", "p2plane"};
    for (int e = 0; e < 2; e++) {
        auto ref = open3d::RegistrationICP(source, target, max_d, init, *ests[e]);
        auto gpu = visma_b200::RegistrationICP(source, target, max_d, init, *ests[e]);
        out("\"%s\": {\"dT\": %.3e, \"fitness_ref\": %.17g, \"fitness_gpu\": %.17g, \"rmse_ref\": %.17g, "
               "\"rmse_gpu\": %.17g, \"ncorr_ref\": %zu, \"ncorr_gpu\": %zu}, ",
               names[e], max_abs_diff(ref.transformation_, gpu.transformation_), ref.fitness_, gpu.fitness_,
               ref.inlier_rmse_, gpu.inlier_rmse_, ref.correspondence_set_.size(), gpu.correspondence_set_.size());
    }
    // (2) plumbing (BASELINE config 1): the reference's own CPU loop driving the GPU estimator plug-in
    {
        open3d::cicp::TransformationEstimationPointToPoint4DoFB200 gpu_est;
        auto ref = open3d::RegistrationICP(source, target, max_d, init, p2p);
        auto mix = open3d::RegistrationICP(source, target, max_d, init, gpu_est);
        out("\"cicp_plugin\": {\"dT\": %.3e, \"ncorr_ref\": %zu, \"ncorr_mix\": %zu}, ",
               max_abs_diff(ref.transformation_, mix.transformation_), ref.correspondence_set_.size(),
               mix.correspondence_set_.size());
    }
    // (2b) the gravity-constrained 4-DoF estimator class: the reference's CPU loop driving the GPU estimator vs
    // the whole loop on the GPU; and cicp::...4DoF::ComputeRMSE (src/constrained_ICP.cpp:13-23) on the GPU
    {
        visma_b200::TransformationEstimationPointToPlaneGravity grav(Eigen::Vector3d(0, 1, 0));
        auto mix = open3d::RegistrationICP(source, target, max_d, init, grav);
        auto gpu = visma_b200::RegistrationICP(source, target, max_d, init, grav);
        Eigen::Matrix4d d = gpu.transformation_ * init.inverse();
        out("\"gravity\": {\"dT\": %.3e, \"ncorr_mix\": %zu, \"ncorr_gpu\": %zu, \"roll_pitch\": %.3e}, ",
            max_abs_diff(mix.transformation_, gpu.transformation_), mix.correspondence_set_.size(),
            gpu.correspondence_set_.size(), std::abs(d(1, 1) - 1.0) + std::abs(d(0, 1)) + std::abs(d(2, 1)));
        open3d::cicp::TransformationEstimationPointToPoint4DoFB200 gpu_est;
        auto ref = open3d::RegistrationICP(source, target, max_d, init, p2p);
        open3d::PointCloud moved = source;
        moved.Transform(ref.transformation_);
        const double r_ref = p2p.ComputeRMSE(moved, target, ref.correspondence_set_);
        const double r_gpu = gpu_est.ComputeRMSE(moved, target, ref.correspondence_set_);
        out("\"rmse\": {\"ref\": %.17g, \"gpu\": %.17g, \"empty\": %g}, ", r_ref, r_gpu,
            gpu_est.ComputeRMSE(moved, target, open3d::CorrespondenceSet()));
    }
    // (2c) a batch mixing sources with and without normals under point-to-plane: the reference decides per call
    // (Registration.cpp:152-157), so only the bare source returns RegistrationResult(init)
    {
        open3d::PointCloud bare, none;
        bare.points_ = source.points_;
        visma_b200::Scene scene(target, max_d);
        auto res = visma_b200::RegistrationICPBatch({&source, &bare, &none, &source}, scene, max_d,
                                                    {init, init, init, init}, p2l);
        auto solo = visma_b200::RegistrationICP(source, target, max_d, init, p2l);
        out("\"mixed_normals\": {\"with_dT\": %.3e, \"with2_dT\": %.3e, \"bare_dT\": %.3e, \"bare_fitness\": %g, "
            "\"empty_dT\": %.3e, \"with_ncorr\": %zu}, ",
            max_abs_diff(res[0].transformation_, solo.transformation_),
            max_abs_diff(res[3].transformation_, solo.transformation_), max_abs_diff(res[1].transformation_, init),
            res[1].fitness_, max_abs_diff(res[2].transformation_, init), res[0].correspondence_set_.size());
    }
    // (3) error behaviour: invalid distance and missing normals return RegistrationResult(init)
    {
        auto a = visma_b200::RegistrationICP(source, target, -1.0, init, p2p);
        open3d::PointCloud bare;
        bare.points_ = source.points_;
        auto b = visma_b200::RegistrationICP(bare, target, max_d, init, p2l);
        out("\"errors\": {\"bad_distance_dT\": %.3e, \"no_normals_dT\": %.3e, \"no_normals_fitness\": %g}, ",
               max_abs_diff(a.transformation_, init), max_abs_diff(b.transformation_, init), b.fitness_);
    }
    // (4) VoxelDownSample: same point set as the reference (order differs: unordered_map vs voxel index)
    {
        auto ref = open3d::VoxelDownSample(target, 0.05);
        auto gpu = visma_b200::VoxelDownSample(target, 0.05);
        double sr[3] = {0, 0, 0}, sg[3] = {0, 0, 0};
        for (auto &p : ref->points_) for (int a = 0; a < 3; a++) sr[a] += p[a];
        for (auto &p : gpu->points_) for (int a = 0; a < 3; a++) sg[a] += p[a];
        out("\"voxel\": {\"n_ref\": %zu, \"n_gpu\": %zu, \"dsum\": %.3e}, ", ref->points_.size(),
               gpu->points_.size(), std::abs(sr[0] - sg[0]) + std::abs(sr[1] - sg[1]) + std::abs(sr[2] - sg[2]));
    }
    // (5) the Renderer class compiles with the reference's call sequence (render/tools/render_depth.cpp:30-47)
    {
        visma_b200::Renderer ren(480, 640, 3, 3);
        ren.SetCamera(0.05f, 10.0f, 400.f, 400.f, 320.f, 400.f);  // the tool passes fy as cy (:31)
        ren.SetCamera(visma_b200::Renderer::Mat4fc::Identity());
        std::vector<float> V = {-0.5f, -0.5f, 0.f, 0.5f, -0.5f, 0.f, 0.f, 0.5f, 0.f};
        std::vector<int> F = {0, 1, 2};
        ren.SetMesh(V, F);
        visma_b200::Renderer::Mat4fc model = visma_b200::Renderer::Mat4fc::Identity();
        model(2, 3) = 1.0f;
        std::vector<float> depth(480 * 640);
        ren.RenderDepth(model, depth.data());
        int covered = 0;
        float zmin = 1.f;
        for (float z : depth) if (z < 1.f) { covered++; zmin = std::min(zmin, z); }
        out("\"render\": {\"covered\": %d, \"z_lin\": %.6f}", covered,
               visma_b200::LinearizeDepth<float>(zmin, 0.05f, 10.0f));
    }
    out("}");
    printf("\nJSON:%s\n", g_json.c_str());
    return 0;
}
