// voxel.cu — vb200_voxel_downsample: replaces open3d::VoxelDownSample
// (O3D/src/Core/Geometry/DownSample.cpp:179-220, AccumulatedPoint :38-87), the single-threaded
// unordered_map pass that precedes every ICP in VISMA (src/evaluation.cpp:258, src/annotation.cpp:112).
//
// The reference's voxel index space is sparse (up to INT_MAX per axis), so voxels are found with an
// open-addressing hash table on the packed 3 x 21-bit index (atomicCAS insert).  Points are then counting-
// sorted by table slot; one thread per occupied slot puts its point indices in ascending order and
// accumulates them sequentially — the order in which the reference's single loop over i adds them — so the
// averages are bit-identical to the CPU result.  Output is ordered by each voxel's first point in the
// input (a prefix sum over "is the first point of its voxel" flags), which is deterministic; the
// reference's order is the unordered_map's.
#include <limits.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "sort.cuh"

namespace vb {

namespace {

constexpr unsigned long long kEmpty = ~0ull;

struct VoxParams {
    double mn[3];
    double voxel;
};

__device__ __forceinline__ unsigned long long voxel_key(const VoxParams &vp, const double *p) {
    // ref_coord = (p - voxel_min_bound) / voxel_size; index = int(floor(ref_coord))  (DownSample.cpp:200-203)
    unsigned long long ix = (unsigned long long)(long long)floor(__ddiv_rn(__dsub_rn(p[0], vp.mn[0]), vp.voxel));
    unsigned long long iy = (unsigned long long)(long long)floor(__ddiv_rn(__dsub_rn(p[1], vp.mn[1]), vp.voxel));
    unsigned long long iz = (unsigned long long)(long long)floor(__ddiv_rn(__dsub_rn(p[2], vp.mn[2]), vp.voxel));
    return (iz << 42) | (iy << 21) | ix;  // each < 2^21, checked on the host
}

__device__ __forceinline__ unsigned hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__global__ void __launch_bounds__(256) k_vox_insert(VoxParams vp, const double *__restrict__ xyz, int64_t n,
                                                    unsigned long long *__restrict__ table, unsigned mask,
                                                    int *__restrict__ pslot, int *__restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = voxel_key(vp, xyz + 3 * i);
    unsigned slot = hash64(key) & mask;
    for (;;) {
        unsigned long long prev = atomicCAS(table + slot, kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        slot = (slot + 1) & mask;
    }
    pslot[i] = (int)slot;
    atomicAdd(counts + slot, 1);
}

__global__ void __launch_bounds__(256) k_vox_scatter(const int *__restrict__ pslot, int64_t n,
                                                     const int *__restrict__ start, int *__restrict__ cursor,
                                                     int *__restrict__ sidx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = pslot[i];
    sidx[start[k] + atomicAdd(cursor + k, 1)] = (int)i;
}

// one thread per slot: order the slot's points by index, flag the first one
__global__ void __launch_bounds__(128) k_vox_sort(int64_t nslot, const int *__restrict__ start,
                                                  int *__restrict__ sidx, int *__restrict__ first_flag) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nslot) return;
    int s0 = start[c], s1 = start[c + 1];
    if (s1 <= s0) return;
    cell_sort(sidx + s0, s1 - s0);
    first_flag[sidx[s0]] = 1;
}

__global__ void __launch_bounds__(128) k_vox_reduce(int64_t nslot, const int *__restrict__ start,
                                                    const int *__restrict__ rank, const int *__restrict__ sidx,
                                                    const double *__restrict__ xyz, const double *__restrict__ nrm,
                                                    double *__restrict__ out_xyz, double *__restrict__ out_nrm) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nslot) return;
    int s0 = start[c], s1 = start[c + 1];
    if (s1 <= s0) return;
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
    const int o = rank[sidx[s0]];  // output slot = number of voxels whose first point comes earlier
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
    if (n < 0 || n > 0x3fffffff || (n > 0 && (!xyz || !out_xyz))) return VB200_ERR_INVALID;
    if (!(voxel_size > 0.0)) return VB200_ERR_INVALID;  // reference: empty output (DownSample.cpp:183-186)
    if (n == 0) return VB200_OK;
    VB_TRY(select_device(device));
    cudaStream_t st;
    VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    DevBuf<double> d_xyz(st), d_nrm(st), d_out(st), d_out_n(st);  // stream-ordered pool: no cudaMalloc / cudaFree per call
    VB_CUDA(d_xyz.alloc(3 * (size_t)n));
    VB_CUDA(h2d_async(d_xyz.p, xyz, sizeof(double) * 3 * (size_t)n, st));
    if (nrm) {
        VB_CUDA(d_nrm.alloc(3 * (size_t)n));
        VB_CUDA(h2d_async(d_nrm.p, nrm, sizeof(double) * 3 * (size_t)n, st));
    }
    double lo[3], hi[3];
    VB_TRY(device_bbox(d_xyz.p, n, lo, hi, st));
    VoxParams vp;
    vp.voxel = voxel_size;
    double ext = 0.0;
    for (int a = 0; a < 3; a++) {
        vp.mn[a] = lo[a] - voxel_size * 0.5;  // DownSample.cpp:189
        double mx = hi[a] + voxel_size * 0.5;  // :190
        ext = std::max(ext, mx - vp.mn[a]);
    }
    if (voxel_size * (double)INT_MAX < ext) return VB200_ERR_INVALID;  // "voxel_size is too small" (:191-195)
    if (ext / voxel_size >= 2097152.0) return VB200_ERR_INVALID;       // 21 bits per axis in the packed key
    unsigned nslot = 1024;
    while ((int64_t)nslot < 2 * n) nslot <<= 1;
    DevBuf<unsigned long long> d_table(st);
    DevBuf<int> d_pslot(st), d_counts(st), d_start(st), d_first(st), d_rank(st), d_sidx(st), d_total(st);
    VB_CUDA(d_table.alloc(nslot));
    VB_CUDA(d_pslot.alloc((size_t)n));
    VB_CUDA(d_sidx.alloc((size_t)n));
    VB_CUDA(d_counts.alloc((size_t)nslot + 1));
    VB_CUDA(d_start.alloc((size_t)nslot + 1));
    VB_CUDA(d_first.alloc((size_t)n + 1));
    VB_CUDA(d_rank.alloc((size_t)n + 1));
    VB_CUDA(d_total.alloc(1));
    VB_CUDA(cudaMemsetAsync(d_table.p, 0xff, sizeof(unsigned long long) * nslot, st));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * ((size_t)nslot + 1), st));
    VB_CUDA(cudaMemsetAsync(d_first.p, 0, sizeof(int) * ((size_t)n + 1), st));
    k_vox_insert<<<div_up(n, 256), 256, 0, st>>>(vp, d_xyz.p, n, d_table.p, nslot - 1, d_pslot.p, d_counts.p);
    VB_TRY(exclusive_scan_i32(d_counts.p, d_start.p, (int64_t)nslot + 1, nullptr, st));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * ((size_t)nslot + 1), st));
    k_vox_scatter<<<div_up(n, 256), 256, 0, st>>>(d_pslot.p, n, d_start.p, d_counts.p, d_sidx.p);
    k_vox_sort<<<div_up(nslot, 128), 128, 0, st>>>(nslot, d_start.p, d_sidx.p, d_first.p);
    VB_TRY(exclusive_scan_i32(d_first.p, d_rank.p, n + 1, d_total.p, st));
    int nvox = 0;
    VB_CUDA(cudaMemcpyAsync(&nvox, d_total.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    VB_CUDA(d_out.alloc(3 * (size_t)nvox));
    if (nrm && out_nrm) VB_CUDA(d_out_n.alloc(3 * (size_t)nvox));
    k_vox_reduce<<<div_up(nslot, 128), 128, 0, st>>>(nslot, d_start.p, d_rank.p, d_sidx.p, d_xyz.p, d_nrm.p,
                                                     d_out.p, d_out_n.p);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(out_xyz, d_out.p, sizeof(double) * 3 * (size_t)nvox, cudaMemcpyDeviceToHost, st));
    if (nrm && out_nrm)
        VB_CUDA(cudaMemcpyAsync(out_nrm, d_out_n.p, sizeof(double) * 3 * (size_t)nvox, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    *out_n = nvox;
    return VB200_OK;
}
