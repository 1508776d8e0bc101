// sharded_b200.h — the multi-GPU form of VISMA's per-object alignment loop (src/annotation.cpp:103-141) for C++
// hosts: one host thread or process per GPU, every rank holds a replica of the scene, rank r of `world` aligns the
// objects {b : b mod world == r} with one batched launch sequence, and ONE ncclAllGather of the pose table (20
// doubles per object: the 4x4 row-major, fitness, inlier rmse, correspondence count, iterations) leaves all the
// results on every rank.  There is no per-iteration exchange: the objects are independent problems.
// RegistrationICPGlobalSharded is the other split — ONE cloud sharded over the ranks, one 256-byte ncclAllReduce per
// iteration (ICPRefinement's global transform).
//
// Header-only on top of registration_b200.h; NCCL and the CUDA runtime are dependencies of the INCLUDING target
// only (libvisma_b200.so itself links neither NCCL nor MPI).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include "registration_b200.h"

namespace visma_b200 {

constexpr int kPoseRow = 20;  // T (16, row-major) + fitness + rmse + ncorr + iterations

inline std::vector<int> ShardObjects(int n_objects, int rank, int world) {
    std::vector<int> mine;
    for (int b = rank; b < n_objects; b += world) mine.push_back(b);
    return mine;
}
inline int RowsPerRank(int n_objects, int world) { return (n_objects + world - 1) / world; }

/// All `sources.size()` problems, sharded over the ranks of `comm`.  Every rank passes the same global lists and its
/// own scene replica (on the device the communicator's rank is bound to).  Returns the result of EVERY object on
/// every rank; correspondence sets are filled for the rank's own objects only (the others carry their size in
/// (*ncorr_all)[b] when ncorr_all is given).  Errors follow RegistrationICPBatch: a failed rank contributes
/// RegistrationResult(init) rows, and a failed collective returns the inits.
inline std::vector<open3d::RegistrationResult> RegistrationICPSharded(
        const std::vector<const open3d::PointCloud *> &sources, const Scene &target,
        double max_correspondence_distance, const std::vector<Eigen::Matrix4d> &inits,
        const open3d::TransformationEstimation &estimation, const open3d::ICPConvergenceCriteria &criteria,
        ncclComm_t comm, int rank, int world, cudaStream_t stream = nullptr,
        std::vector<int> *ncorr_all = nullptr) {
    const int n = (int)sources.size();
    std::vector<open3d::RegistrationResult> out;
    for (int b = 0; b < n; b++) out.emplace_back(inits[b]);
    if (ncorr_all) ncorr_all->assign(n, 0);
    if (n == 0) return out;
    const std::vector<int> mine = ShardObjects(n, rank, world);
    std::vector<const open3d::PointCloud *> my_src;
    std::vector<Eigen::Matrix4d> my_init;
    for (int b : mine) { my_src.push_back(sources[b]); my_init.push_back(inits[b]); }
    std::vector<open3d::RegistrationResult> my_res =
            RegistrationICPBatch(my_src, target, max_correspondence_distance, my_init, estimation, criteria);
    // this rank's rows of the table (padded to the same size on every rank)
    const int rows = RowsPerRank(n, world);
    std::vector<double> send((size_t)rows * kPoseRow, 0.0), recv((size_t)world * rows * kPoseRow, 0.0);
    for (size_t k = 0; k < mine.size(); k++) {
        double *r = send.data() + k * kPoseRow;
        ToRowMajor(my_res[k].transformation_, r);
        r[16] = my_res[k].fitness_;
        r[17] = my_res[k].inlier_rmse_;
        r[18] = (double)my_res[k].correspondence_set_.size();
        r[19] = 0.0;  // (iterations are not part of open3d::RegistrationResult)
    }
    double *d_send = nullptr, *d_recv = nullptr;
    bool ok = cudaMalloc((void **)&d_send, send.size() * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&d_recv, recv.size() * sizeof(double)) == cudaSuccess &&
              cudaMemcpyAsync(d_send, send.data(), send.size() * sizeof(double), cudaMemcpyHostToDevice, stream) == cudaSuccess &&
              ncclAllGather(d_send, d_recv, send.size(), ncclDouble, comm, stream) == ncclSuccess &&
              cudaMemcpyAsync(recv.data(), d_recv, recv.size() * sizeof(double), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
              cudaStreamSynchronize(stream) == cudaSuccess;
    cudaFree(d_send);
    cudaFree(d_recv);
    if (!ok) {
        open3d::PrintError("visma_b200::RegistrationICPSharded: all-gather of the pose table failed\n");
        return out;
    }
    for (int r = 0; r < world; r++) {
        int k = 0;
        for (int b = r; b < n; b += world, k++) {
            const double *row = recv.data() + ((size_t)r * rows + k) * kPoseRow;
            out[b].transformation_ = FromRowMajor(row);
            out[b].fitness_ = row[16];
            out[b].inlier_rmse_ = row[17];
            if (ncorr_all) (*ncorr_all)[b] = (int)row[18];
        }
    }
    for (size_t k = 0; k < mine.size(); k++) out[mine[k]].correspondence_set_ = std::move(my_res[k].correspondence_set_);
    return out;
}

/// ONE source cloud sharded over the ranks of `comm` — feh::ICPRefinement's single global transform
/// (src/evaluation.cpp:244-274) at multi-GPU scale.  Every rank passes ITS slice of the cloud and its own scene
/// replica; an iteration is vb200_batch_pass (correspondence pass over the slice, 32 totals left on the device) ->
/// ncclAllReduce of those 256 bytes on the library's stream -> vb200_batch_solve (fitness / rmse / convergence test /
/// estimator update from the combined totals).  Every rank sees bit-identical totals and applies the identical
/// update, so the transforms stay consistent without a broadcast; the result (without a correspondence set: it
/// would be the slice's) is the same on every rank.  Errors follow RegistrationICP: RegistrationResult(init).
inline open3d::RegistrationResult RegistrationICPGlobalSharded(
        const open3d::PointCloud &source_slice, const Scene &target, double max_correspondence_distance,
        const Eigen::Matrix4d &init, const open3d::TransformationEstimation &estimation,
        const open3d::ICPConvergenceCriteria &criteria, ncclComm_t comm) {
    open3d::RegistrationResult out(init);
    if (max_correspondence_distance <= 0.0) {
        open3d::PrintError("Error: Invalid max_correspondence_distance.\n");
        return out;
    }
    const int kind = EstimatorKind(estimation);
    // (every rank must take the same branch: a slice may be empty, so the cloud-level property is the caller's to keep)
    if (kind != VB200_EST_P2P && !source_slice.points_.empty() && !source_slice.HasNormals()) {
        open3d::PrintError("Error: TransformationEstimationPointToPlane requires pre-computed normal vectors.\n");
        return out;
    }
    cudaStream_t st = (cudaStream_t)vb200_scene_stream(target.handle());
    const int64_t off[2] = {0, (int64_t)source_slice.points_.size()};
    const double dummy[3] = {0.0, 0.0, 0.0};
    const double *xyz = source_slice.points_.empty() ? dummy : Raw(source_slice.points_);
    vb200_batch_t *batch = nullptr;
    double T0[16], T[16], fit = 0.0, rmse = 0.0;
    ToRowMajor(init, T0);
    double *d_totals = nullptr;       // 32 totals, then the slice's point count as a double (exact below 2^53)
    bool ok = vb200_batch_create(target.handle(), xyz, kind != VB200_EST_P2P ? xyz : nullptr, off, 1, &batch) == VB200_OK &&
              vb200_batch_set_problems(batch, nullptr, T0, 1) == VB200_OK &&
              cudaMalloc((void **)&d_totals, 33 * sizeof(double)) == cudaSuccess &&
              cudaMemsetAsync(d_totals, 0, 33 * sizeof(double), st) == cudaSuccess &&
              vb200_batch_set_totals_buffer(batch, d_totals) == VB200_OK;
    int64_t n_global = 0;
    if (ok) {
        double n_local = (double)off[1], n_sum = 0.0;
        ok = cudaMemcpyAsync(d_totals + 32, &n_local, sizeof(double), cudaMemcpyHostToDevice, st) == cudaSuccess &&
             ncclAllReduce(d_totals + 32, d_totals + 32, 1, ncclDouble, ncclSum, comm, st) == ncclSuccess &&
             cudaMemcpyAsync(&n_sum, d_totals + 32, sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
        n_global = (int64_t)n_sum;
    }
    for (int it = 0; ok && it <= criteria.max_iteration_; it++) {
        ok = vb200_batch_pass(batch, kind, max_correspondence_distance) == VB200_OK &&
             ncclAllReduce(d_totals, d_totals, 32, ncclDouble, ncclSum, comm, st) == ncclSuccess &&
             vb200_batch_solve(batch, kind, GravityAxis(estimation), max_correspondence_distance,
                               criteria.relative_fitness_, criteria.relative_rmse_, criteria.max_iteration_, it,
                               &n_global) == VB200_OK;
    }
    int32_t nc = 0, iters = 0;
    ok = ok && vb200_batch_results(batch, T, &fit, &rmse, &nc, &iters) == VB200_OK;
    if (batch) vb200_batch_destroy(batch);
    cudaFree(d_totals);
    if (!ok) {
        open3d::PrintError("visma_b200::RegistrationICPGlobalSharded: %s\n", vb200_last_error());
        return out;
    }
    out.transformation_ = FromRowMajor(T);
    out.fitness_ = fit;
    out.inlier_rmse_ = rmse;
    return out;
}

}  // namespace visma_b200
