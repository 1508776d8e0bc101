// knn.cu — vb200_knn1 / vb200_knn1_device: radius-bounded 1-NN for a batch of query points.
// Replaces the per-point KDTreeFlann::SearchHybrid(query, radius, 1) loop of
// GetRegistrationResultAndCorrespondences (O3D/src/Core/Registration/Registration.cpp:62-72,
// O3D/src/Core/Geometry/KDTreeFlann.cpp:165-189).  Results are bit-identical to the reference's double
// arithmetic: same neighbour index, same d2 (see grid.cuh).
#include <math.h>

#include "scene.cuh"

namespace vb {

namespace {

constexpr int kTpb = 256;

__global__ void __launch_bounds__(kTpb) k_knn1(GridDev G, const double *__restrict__ q, int64_t nq,
                                               const int *__restrict__ perm, double r2, float r2_ub,
                                               int *__restrict__ out_idx, double *__restrict__ out_d2) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < nq;  // whole warps stay alive: the search is warp-cooperative
    const int64_t i = live ? perm[s] : 0;  // queries are visited in grid order, results land in caller order
    double x = 0.0, y = 0.0, z = 0.0, d2 = 0.0;
    QueryCtx c;
    bool inside = false;
    if (live) {
        x = q[3 * i]; y = q[3 * i + 1]; z = q[3 * i + 2];
        inside = make_query(G.p, x, y, z, c);
    }
    __shared__ LaneRuns<kTpb> runs;
    const int bs = nn_search_hybrid<kTpb>(G, inside, c, x, y, z, r2, r2_ub, -1, runs, &d2);
    if (live) {
        out_idx[i] = bs >= 0 ? __ldg(G.orig + bs) : -1;
        out_d2[i] = bs >= 0 ? d2 : 0.0;
    }
}


// one warp per query (few / scattered queries)
__global__ void __launch_bounds__(kTpb) k_knn1_wpq(GridDev G, const double *__restrict__ q, int64_t nq, double r2,
                                                   float r2_ub, int *__restrict__ out_idx,
                                                   double *__restrict__ out_d2) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq) return;  // warp-uniform exit
    const double x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    QueryCtx c;
    int bs = -1;
    double d2 = 0.0;
    if (make_query(G.p, x, y, z, c)) bs = nn_search_wpq(G, c, x, y, z, r2, r2_ub, &d2);
    if ((threadIdx.x & 31) == 0) {
        out_idx[i] = bs >= 0 ? __ldg(G.orig + bs) : -1;
        out_d2[i] = bs >= 0 ? d2 : 0.0;
    }
}

constexpr int64_t kWarpPerQueryMax = 262144;  // below this many queries a warp per query fills the GPU better

int knn1_launch(Scene *sc, const double *d_q, int64_t nq, double radius, int *d_idx, double *d_d2) {
    if (!(radius > 0.0) || radius > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (nq == 0) return VB200_OK;
    if (nq > 0x7fffffff) return VB200_ERR_INVALID;
    const double r2 = (double)(float)(radius * radius);  // KDTreeFlann.cpp:185
    if (nq <= kWarpPerQueryMax) {
        k_knn1_wpq<<<div_up(nq * 32, kTpb), kTpb, 0, sc->stream>>>(sc->grid, d_q, nq, r2,
                                                                   r2_upper_bound(sc->grid.p, r2), d_idx, d_d2);
        VB_CUDA(cudaGetLastError());
        return VB200_OK;
    }
    DevBuf<int> d_perm(sc->stream);
    VB_CUDA(d_perm.alloc((size_t)nq));
    VB_TRY(grid_order_points(sc, d_q, nq, d_perm.p));
    k_knn1<<<div_up(nq, kTpb), kTpb, 0, sc->stream>>>(sc->grid, d_q, nq, d_perm.p, r2,
                                                      r2_upper_bound(sc->grid.p, r2), d_idx, d_d2);
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

}  // namespace

}  // namespace vb

using vb::Scene;

extern "C" int vb200_knn1_device(vb200_scene_t *scene, const void *d_q_xyz, int64_t Q, double radius,
                                 void *d_out_idx, void *d_out_d2) {
    if (!scene || Q < 0 || (Q > 0 && (!d_q_xyz || !d_out_idx || !d_out_d2))) return VB200_ERR_INVALID;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    return vb::knn1_launch(sc, (const double *)d_q_xyz, Q, radius, (int *)d_out_idx, (double *)d_out_d2);
}

extern "C" int vb200_knn1(vb200_scene_t *scene, const double *q_xyz, int64_t Q, double radius, int32_t *out_idx,
                          double *out_d2) {
    if (!scene || Q < 0 || (Q > 0 && (!q_xyz || !out_idx || !out_d2))) return VB200_ERR_INVALID;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    if (!(radius > 0.0) || radius > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (Q == 0) return VB200_OK;
    vb::DevBuf<double> d_q, d_d2;
    vb::DevBuf<int> d_idx;
    VB_CUDA(d_q.alloc(3 * (size_t)Q));
    VB_CUDA(d_d2.alloc((size_t)Q));
    VB_CUDA(d_idx.alloc((size_t)Q));
    VB_CUDA(cudaMemcpyAsync(d_q.p, q_xyz, sizeof(double) * 3 * (size_t)Q, cudaMemcpyHostToDevice, sc->stream));
    VB_TRY(vb::knn1_launch(sc, d_q.p, Q, radius, d_idx.p, d_d2.p));
    VB_CUDA(cudaMemcpyAsync(out_idx, d_idx.p, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost, sc->stream));
    VB_CUDA(cudaMemcpyAsync(out_d2, d_d2.p, sizeof(double) * (size_t)Q, cudaMemcpyDeviceToHost, sc->stream));
    VB_CUDA(cudaStreamSynchronize(sc->stream));
    return VB200_OK;
}
