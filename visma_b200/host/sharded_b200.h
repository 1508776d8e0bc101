// sharded_b200.h — the multi-GPU form of VISMA's per-object alignment loop (src/annotation.cpp:103-141) for C++
// hosts: one host thread or process per GPU, every rank holds a replica of the scene, rank r of `world` aligns the
// objects {b : b mod world == r} with one batched launch sequence, and ONE ncclAllGather of the pose table (20
// doubles per object: the 4x4 row-major, fitness, inlier rmse, correspondence count, iterations) leaves all the
// results on every rank.  There is no per-iteration exchange: the objects are independent problems.
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

}  // namespace visma_b200
