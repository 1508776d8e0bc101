// scene.cuh — the resident target cloud behind vb200_scene_t.
#pragma once

#include "grid.cuh"

namespace vb {

struct Scene {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // created on first use by vb200_icp_run (second half of a pipelined batch)
    GridDev grid{};
    int64_t n = 0;
    int64_t ncoarse = 0;
    int64_t nfine = 0;
    bool has_normals = false;
};

int scene_build(Scene *sc, const double *h_xyz, const double *h_nrm, int64_t n, double max_radius);
void scene_free(Scene *sc);
int grid_order_points(const Scene *sc, const double *d_xyz, int64_t n, int *d_perm);

}  // namespace vb
