// voxel.cu — vb200_voxel_downsample: replaces open3d::VoxelDownSample
// (O3D/src/Core/Geometry/DownSample.cpp:179-220, AccumulatedPoint :38-87), the single-threaded
// unordered_map pass that precedes every ICP in VISMA (src/evaluation.cpp:258, src/annotation.cpp:112).
//
// Dense voxel grid + counting sort: histogram -> scan -> scatter, then one thread per occupied voxel puts
// its point indices in ascending order and accumulates them sequentially — the same order the reference's
// single loop over i adds them — so the averages are bit-identical to the CPU result.  Output is ordered by
// voxel index (z-major); the reference's order is unordered_map iteration order.
#include <limits.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "sort.cuh"

namespace vb {

namespace {

struct VoxParams {
    double mn[3];
    double voxel;
    int dim[3];
};

__device__ __forceinline__ int64_t voxel_key(const VoxParams &vp, const double *p) {
    // ref_coord = (p - voxel_min_bound) / voxel_size; index = int(floor(ref_coord))  (DownSample.cpp:200-203)
    int ix = (int)floor(__ddiv_rn(__dsub_rn(p[0], vp.mn[0]), vp.voxel));
    int iy = (int)floor(__ddiv_rn(__dsub_rn(p[1], vp.mn[1]), vp.voxel));
    int iz = (int)floor(__ddiv_rn(__dsub_rn(p[2], vp.mn[2]), vp.voxel));
    ix = min(max(ix, 0), vp.dim[0] - 1);
    iy = min(max(iy, 0), vp.dim[1] - 1);
    iz = min(max(iz, 0), vp.dim[2] - 1);
    return ((int64_t)iz * vp.dim[1] + iy) * vp.dim[0] + ix;
}

__global__ void __launch_bounds__(256) k_vox_count(VoxParams vp, const double *__restrict__ xyz, int64_t n,
                                                   int *__restrict__ key, int *__restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = (int)voxel_key(vp, xyz + 3 * i);
    key[i] = k;
    atomicAdd(counts + k, 1);
}

__global__ void __launch_bounds__(256) k_vox_scatter(const int *__restrict__ key, int64_t n,
                                                     const int *__restrict__ start, int *__restrict__ cursor,
                                                     int *__restrict__ sidx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = key[i];
    sidx[start[k] + atomicAdd(cursor + k, 1)] = (int)i;
}

__global__ void __launch_bounds__(256) k_vox_flags(const int *__restrict__ start, int64_t ncell,
                                                   int *__restrict__ flag) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncell) flag[c] = start[c + 1] > start[c] ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_vox_reduce(int64_t ncell, const int *__restrict__ start,
                                                    const int *__restrict__ rank, int *__restrict__ sidx,
                                                    const double *__restrict__ xyz, const double *__restrict__ nrm,
                                                    double *__restrict__ out_xyz, double *__restrict__ out_nrm) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int s0 = start[c], s1 = start[c + 1];
    if (s1 <= s0) return;
    cell_sort(sidx + s0, s1 - s0);
    double p[3] = {0, 0, 0}, q[3] = {0, 0, 0};
    for (int s = s0; s < s1; s++) {
        int64_t i = sidx[s];
        p[0] = __dadd_rn(p[0], xyz[3 * i]);       // AddPoint: point_ += cloud.points_[index]
        p[1] = __dadd_rn(p[1], xyz[3 * i + 1]);
        p[2] = __dadd_rn(p[2], xyz[3 * i + 2]);
        if (nrm) {
            double a = nrm[3 * i], b = nrm[3 * i + 1], cc = nrm[3 * i + 2];
            if (!isnan(a) && !isnan(b) && !isnan(cc)) {  // DownSample.cpp:52-56
                q[0] = __dadd_rn(q[0], a);
                q[1] = __dadd_rn(q[1], b);
                q[2] = __dadd_rn(q[2], cc);
            }
        }
    }
    const int o = rank[c];
    const double cnt = (double)(s1 - s0);
    out_xyz[3 * (int64_t)o] = __ddiv_rn(p[0], cnt);  // GetAveragePoint
    out_xyz[3 * (int64_t)o + 1] = __ddiv_rn(p[1], cnt);
    out_xyz[3 * (int64_t)o + 2] = __ddiv_rn(p[2], cnt);
    if (nrm && out_nrm) {
        // normal_.normalized(): v / sqrt(x*x + y*y + z*z), sum left to right
        double l2 = __dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2]));
        double l = __dsqrt_rn(l2);
        for (int a = 0; a < 3; a++) out_nrm[3 * (int64_t)o + a] = l > 0.0 ? __ddiv_rn(q[a], l) : q[a];
    }
}

}  // namespace

}  // namespace vb

extern "C" int vb200_voxel_downsample(const double *xyz, const double *nrm, int64_t n, double voxel_size,
                                      int device, double *out_xyz, double *out_nrm, int64_t *out_n) {
    using namespace vb;
    if (!out_n) return VB200_ERR_INVALID;
    *out_n = 0;
    if (n < 0 || n > 0x7fffffff || (n > 0 && (!xyz || !out_xyz))) return VB200_ERR_INVALID;
    if (!(voxel_size > 0.0)) return VB200_ERR_INVALID;  // reference: empty output (DownSample.cpp:183-186)
    if (n == 0) return VB200_OK;
    VB_TRY(select_device(device));
    cudaStream_t st;
    VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    DevBuf<double> d_xyz, d_nrm, d_out, d_out_n;
    VB_CUDA(d_xyz.alloc(3 * (size_t)n));
    VB_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
    if (nrm) {
        VB_CUDA(d_nrm.alloc(3 * (size_t)n));
        VB_CUDA(cudaMemcpyAsync(d_nrm.p, nrm, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    double lo[3], hi[3];
    VB_TRY(device_bbox(d_xyz.p, n, lo, hi, st));
    VoxParams vp;
    vp.voxel = voxel_size;
    double ext = 0.0, cells = 1.0;
    for (int a = 0; a < 3; a++) {
        vp.mn[a] = lo[a] - voxel_size * 0.5;  // DownSample.cpp:189
        double mx = hi[a] + voxel_size * 0.5;  // :190
        ext = std::max(ext, mx - vp.mn[a]);
        double d = floor((hi[a] - vp.mn[a]) / voxel_size) + 1.0;
        cells *= d;
        vp.dim[a] = (int)std::min(d, 2147483647.0);
    }
    if (voxel_size * (double)INT_MAX < ext) return VB200_ERR_INVALID;  // "voxel_size is too small" (:191-195)
    if (cells > 1073741824.0) return VB200_ERR_NOMEM;  // dense voxel table capped at 2^30 cells
    const int64_t ncell = (int64_t)cells;
    DevBuf<int> d_key, d_counts, d_start, d_flag, d_rank, d_sidx, d_total;
    VB_CUDA(d_key.alloc((size_t)n));
    VB_CUDA(d_sidx.alloc((size_t)n));
    VB_CUDA(d_counts.alloc((size_t)ncell + 1));
    VB_CUDA(d_start.alloc((size_t)ncell + 1));
    VB_CUDA(d_flag.alloc((size_t)ncell + 1));
    VB_CUDA(d_rank.alloc((size_t)ncell + 1));
    VB_CUDA(d_total.alloc(1));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * ((size_t)ncell + 1), st));
    VB_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int) * ((size_t)ncell + 1), st));
    k_vox_count<<<div_up(n, 256), 256, 0, st>>>(vp, d_xyz.p, n, d_key.p, d_counts.p);
    VB_TRY(exclusive_scan_i32(d_counts.p, d_start.p, ncell + 1, nullptr, st));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * ((size_t)ncell + 1), st));
    k_vox_scatter<<<div_up(n, 256), 256, 0, st>>>(d_key.p, n, d_start.p, d_counts.p, d_sidx.p);
    k_vox_flags<<<div_up(ncell, 256), 256, 0, st>>>(d_start.p, ncell, d_flag.p);
    VB_TRY(exclusive_scan_i32(d_flag.p, d_rank.p, ncell + 1, d_total.p, st));
    int nvox = 0;
    VB_CUDA(cudaMemcpyAsync(&nvox, d_total.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    VB_CUDA(d_out.alloc(3 * (size_t)nvox));
    if (nrm && out_nrm) VB_CUDA(d_out_n.alloc(3 * (size_t)nvox));
    k_vox_reduce<<<div_up(ncell, 128), 128, 0, st>>>(ncell, d_start.p, d_rank.p, d_sidx.p, d_xyz.p, d_nrm.p,
                                                      d_out.p, d_out_n.p);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(out_xyz, d_out.p, sizeof(double) * 3 * (size_t)nvox, cudaMemcpyDeviceToHost, st));
    if (nrm && out_nrm)
        VB_CUDA(cudaMemcpyAsync(out_nrm, d_out_n.p, sizeof(double) * 3 * (size_t)nvox, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    *out_n = nvox;
    return VB200_OK;
}
