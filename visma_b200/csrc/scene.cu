// scene.cu — vb200_scene_*: upload a target cloud and build its two-level NN grid on the GPU.
// Replaces KDTreeFlann::SetGeometry / FLANN buildIndex (O3D/src/Core/Geometry/KDTreeFlann.cpp:70-87,
// 191-208), ~1.5 s single-threaded for 2 M points in the reference and re-run per RegistrationICP call.
//
// Build = counting sort by coarse cell (histogram / scan / scatter), then one thread per occupied coarse
// cell orders its points by (fine cell, original index) so the layout — and therefore every later
// floating-point reduction order — is deterministic.  HBM-bound passes over N points.
#include "scene.cuh"
#include "sort.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace vb {

namespace {

constexpr int kTpb = 256;

__device__ __forceinline__ void fine_coords(const GridParams &g, const double *p, int &gx, int &gy, int &gz) {
    // identical expression to make_query() so build and query agree on cell membership
    // fmin/fmax clamp in double first: far-away or non-finite query coordinates must not overflow the int cast
    gx = (int)floor(fmin(fmax((p[0] - g.lo[0]) * g.inv_fine, 0.0), (double)(g.fdim[0] - 1)));
    gy = (int)floor(fmin(fmax((p[1] - g.lo[1]) * g.inv_fine, 0.0), (double)(g.fdim[1] - 1)));
    gz = (int)floor(fmin(fmax((p[2] - g.lo[2]) * g.inv_fine, 0.0), (double)(g.fdim[2] - 1)));
    gx = min(max(gx, 0), g.fdim[0] - 1);
    gy = min(max(gy, 0), g.fdim[1] - 1);
    gz = min(max(gz, 0), g.fdim[2] - 1);
}

__global__ void __launch_bounds__(kTpb) k_keys(GridParams g, const double *__restrict__ xyz, int64_t n,
                                               int *__restrict__ ckey, unsigned char *__restrict__ fbit,
                                               int *__restrict__ ccount) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int gx, gy, gz;
    fine_coords(g, xyz + 3 * i, gx, gy, gz);
    int cid = ((gz >> 2) * g.cdim[1] + (gy >> 2)) * g.cdim[0] + (gx >> 2);
    ckey[i] = cid;
    fbit[i] = (unsigned char)((gx & 3) + 4 * (gy & 3) + 16 * (gz & 3));
    atomicAdd(ccount + cid, 1);
}

__global__ void __launch_bounds__(kTpb) k_scatter(const int *__restrict__ ckey,
                                                  const unsigned char *__restrict__ fbit, int64_t n,
                                                  const int *__restrict__ cstart, int *__restrict__ cursor,
                                                  unsigned long long *__restrict__ skey) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cid = ckey[i];
    int pos = cstart[cid] + atomicAdd(cursor + cid, 1);
    skey[pos] = ((unsigned long long)fbit[i] << 32) | (unsigned long long)(unsigned int)i;
}

// one WARP per coarse cell: order its keys by (fine cell, index), OR the fine-cell bits into the occupancy mask
__global__ void __launch_bounds__(256) k_sort_cells(int ncoarse, const int *__restrict__ cstart,
                                                    unsigned long long *__restrict__ skey,
                                                    unsigned long long *__restrict__ cmask,
                                                    int *__restrict__ cfcount) {
    const int c = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= ncoarse) return;  // warp-uniform
    const int s0 = cstart[c], s1 = cstart[c + 1];
    unsigned lo = 0u, hi = 0u;
    if (s1 > s0) {
        warp_cell_sort<unsigned long long, 16>(skey + s0, s1 - s0);
        for (int s = s0 + lane; s < s1; s += 32) {
            const unsigned b = (unsigned)(skey[s] >> 32);
            if (b < 32u) lo |= 1u << b; else hi |= 1u << (b - 32u);
        }
    }
    lo = __reduce_or_sync(0xffffffffu, lo);
    hi = __reduce_or_sync(0xffffffffu, hi);
    if (lane == 0) {
        const unsigned long long m = ((unsigned long long)hi << 32) | lo;
        cmask[c] = m;
        cfcount[c] = __popcll(m);
    }
}

__global__ void __launch_bounds__(128) k_fine_starts(int ncoarse, const int *__restrict__ cstart,
                                                     const unsigned long long *__restrict__ skey,
                                                     const unsigned long long *__restrict__ cmask,
                                                     const int *__restrict__ cbase,
                                                     CoarseCell *__restrict__ coarse, int *__restrict__ fstart,
                                                     int nfine_total_slot, int n) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) fstart[nfine_total_slot] = n;
    if (c >= ncoarse) return;
    CoarseCell cc;
    cc.mask = cmask[c];
    cc.base = cbase[c];
    cc.pad = 0;
    coarse[c] = cc;
    int s0 = cstart[c], s1 = cstart[c + 1];
    int prev = -1, k = cc.base;
    for (int s = s0; s < s1; s++) {
        int b = (int)(skey[s] >> 32);
        if (b != prev) { fstart[k++] = s; prev = b; }
    }
}

__global__ void __launch_bounds__(kTpb) k_gather(GridParams g, const unsigned long long *__restrict__ skey,
                                                 int64_t n, const double *__restrict__ xyz_in,
                                                 const double *__restrict__ nrm_in, float4 *__restrict__ hi,
                                                 double *__restrict__ xyz, double *__restrict__ nrm,
                                                 int *__restrict__ orig) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = (int)(unsigned int)skey[s];
    double x = xyz_in[3 * (int64_t)i], y = xyz_in[3 * (int64_t)i + 1], z = xyz_in[3 * (int64_t)i + 2];
    // 32-byte records, written as two 16-byte stores
    double2 *xr = reinterpret_cast<double2 *>(xyz + kPtStride * s);
    xr[0] = make_double2(x, y);
    xr[1] = make_double2(z, 0.0);
    hi[s] = make_float4((float)(x - g.ctr[0]), (float)(y - g.ctr[1]), (float)(z - g.ctr[2]), __int_as_float(i));
    orig[s] = i;
    if (nrm_in) {
        double2 *nr = reinterpret_cast<double2 *>(nrm + kPtStride * s);
        nr[0] = make_double2(nrm_in[3 * (int64_t)i], nrm_in[3 * (int64_t)i + 1]);
        nr[1] = make_double2(nrm_in[3 * (int64_t)i + 2], 0.0);
    }
}

}  // namespace

int scene_build(Scene *sc, const double *h_xyz, const double *h_nrm, int64_t n, double max_radius) {
    cudaStream_t st = sc->stream;
    sc->n = n;
    sc->has_normals = h_nrm != nullptr;
    // raw upload (unsorted), freed after the gather
    // (every temporary and every array of the scene comes from the stream-ordered pool: with the pool's release
    // threshold raised in select_device() a second scene reuses the first one's memory, where cudaMalloc / cudaFree
    // cost milliseconds per call and synchronise the device)
    DevBuf<double> d_in_xyz(st), d_in_nrm(st);
    VB_CUDA(d_in_xyz.alloc(3 * (size_t)n));
    VB_CUDA(h2d_async(d_in_xyz.p, h_xyz, sizeof(double) * 3 * (size_t)n, st));
    if (h_nrm) {
        VB_CUDA(d_in_nrm.alloc(3 * (size_t)n));
        VB_CUDA(h2d_async(d_in_nrm.p, h_nrm, sizeof(double) * 3 * (size_t)n, st));
    }
    // bounding box
    double lo[3], hi[3];
    VB_TRY(device_bbox(d_in_xyz.p, n, lo, hi, st));
    for (int a = 0; a < 3; a++)
        if (!(lo[a] <= hi[a]) || !std::isfinite(lo[a]) || !std::isfinite(hi[a])) return VB200_ERR_INVALID;

    // Coarse cell = scale * max radius (fine cell = 1/4 of it).  scale steps through 1, 1.5, 2, 3, 4 while the
    // occupied fine cells hold fewer than ~5 points on average: the search pays ~40 instructions per visited
    // cell and ~10 per candidate, so nearly-empty cells waste issue slots (measured, profiles/).  The cell is
    // also grown if the dense coarse array would exceed 2^27 cells.
    GridParams &g = sc->grid.p;
    double scale = 1.0;
  for (int attempt = 0;; attempt++) {
    double cell = max_radius * scale;
    for (;;) {
        double cells = 1.0;
        for (int a = 0; a < 3; a++) cells *= floor((hi[a] - lo[a]) / cell) + 3.0;
        if (cells <= 134217728.0) break;
        cell *= 1.25;
    }
    g.cell = cell;
    g.inv_fine = 4.0 / cell;
    g.fine = (float)(cell / 4.0);
    double cmax = 0.0;
    for (int a = 0; a < 3; a++) {
        g.cdim[a] = (int)floor((hi[a] - lo[a]) / cell) + 3;  // one padding cell on each side
        g.fdim[a] = 4 * g.cdim[a];
        g.lo[a] = lo[a] - cell;
        g.ctr[a] = 0.5 * (lo[a] + hi[a]);
        cmax = std::max(cmax, 0.5 * (hi[a] - lo[a]) + 2.0 * cell);
    }
    // error band of the f32 screening distance (see DESIGN.md "exactness"): per-axis difference error
    // e <= u(|q|+|t|+|d|) + centring error, u = 2^-24; |d32 - d2| <= 2*sqrt(3)*e*d + 3e^2 + 4u*d32.
    {
        const double u = 5.9604644775390625e-8;
        double e = u * (2.0 * cmax + cell) + 4e-16 * (cmax + fabs(g.ctr[0]) + fabs(g.ctr[1]) + fabs(g.ctr[2]));
        g.band_a = (float)(2.0 * 1.7320508075688772 * e * 1.01);
        g.band_b = (float)(3.0 * e * e * 1.01 + 1e-30);
        g.band_rel = (float)(4.0 * u * 1.01);
    }
    const int64_t ncoarse = (int64_t)g.cdim[0] * g.cdim[1] * g.cdim[2];
    sc->ncoarse = ncoarse;

    DevBuf<int> d_ckey(st), d_ccount(st), d_cstart(st), d_cfcount(st), d_cbase(st), d_total(st);
    DevBuf<unsigned char> d_fbit(st);
    DevBuf<unsigned long long> d_skey(st), d_cmask(st);
    VB_CUDA(d_ckey.alloc((size_t)n));
    VB_CUDA(d_fbit.alloc((size_t)n));
    VB_CUDA(d_ccount.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cstart.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cfcount.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cbase.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cmask.alloc((size_t)ncoarse));
    VB_CUDA(d_skey.alloc((size_t)n));
    VB_CUDA(d_total.alloc(1));
    VB_CUDA(cudaMemsetAsync(d_ccount.p, 0, sizeof(int) * ((size_t)ncoarse + 1), st));
    VB_CUDA(cudaMemsetAsync(d_cfcount.p, 0, sizeof(int) * ((size_t)ncoarse + 1), st));

    k_keys<<<div_up(n, kTpb), kTpb, 0, st>>>(g, d_in_xyz.p, n, d_ckey.p, d_fbit.p, d_ccount.p);
    VB_CUDA(cudaGetLastError());
    VB_TRY(exclusive_scan_i32(d_ccount.p, d_cstart.p, ncoarse + 1, nullptr, st));
    VB_CUDA(cudaMemsetAsync(d_ccount.p, 0, sizeof(int) * ((size_t)ncoarse + 1), st));  // reuse as cursors
    k_scatter<<<div_up(n, kTpb), kTpb, 0, st>>>(d_ckey.p, d_fbit.p, n, d_cstart.p, d_ccount.p, d_skey.p);
    VB_CUDA(cudaGetLastError());
    k_sort_cells<<<div_up(ncoarse * 32, 256), 256, 0, st>>>((int)ncoarse, d_cstart.p, d_skey.p, d_cmask.p,
                                                             d_cfcount.p);
    VB_CUDA(cudaGetLastError());
    VB_TRY(exclusive_scan_i32(d_cfcount.p, d_cbase.p, ncoarse + 1, d_total.p, st));
    int nfine = 0;
    VB_CUDA(cudaMemcpyAsync(&nfine, d_total.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    sc->nfine = nfine;
    constexpr double occ_target = 5.0;  // points per occupied fine cell below which the cell is grown
    if ((double)n / (double)std::max(nfine, 1) < occ_target && attempt < 4 && n > 1000) {
        static const double kScales[5] = {1.0, 1.5, 2.0, 3.0, 4.0};  // measured optimum is flat between 1.5 and 2
        scale = kScales[attempt + 1];
        continue;  // the DevBufs of this attempt are released by their destructors
    }

    DevBuf<CoarseCell> d_coarse(st);
    DevBuf<int> d_fstart(st), d_orig(st);
    DevBuf<float4> d_hi(st);
    DevBuf<double> d_xyz(st), d_nrm(st);
    VB_CUDA(d_coarse.alloc((size_t)ncoarse));
    VB_CUDA(d_fstart.alloc((size_t)nfine + 1));
    VB_CUDA(d_hi.alloc((size_t)n));
    VB_CUDA(d_xyz.alloc(kPtStride * (size_t)n));
    VB_CUDA(d_orig.alloc((size_t)n));
    if (h_nrm) VB_CUDA(d_nrm.alloc(kPtStride * (size_t)n));
    double *const nrm_out = d_nrm.p;
    k_fine_starts<<<div_up(ncoarse, 128), 128, 0, st>>>((int)ncoarse, d_cstart.p, d_skey.p, d_cmask.p,
                                                         d_cbase.p, d_coarse.p, d_fstart.p, nfine, (int)n);
    VB_CUDA(cudaGetLastError());
    k_gather<<<div_up(n, kTpb), kTpb, 0, st>>>(g, d_skey.p, n, d_in_xyz.p, d_in_nrm.p, d_hi.p, d_xyz.p,
                                               nrm_out, d_orig.p);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaStreamSynchronize(st));

    sc->grid.coarse = d_coarse.take();
    sc->grid.fstart = d_fstart.take();
    sc->grid.hi = d_hi.take();
    sc->grid.xyz = d_xyz.take();
    sc->grid.nrm = h_nrm ? d_nrm.take() : nullptr;
    sc->grid.orig = d_orig.take();
    sc->grid.n = n;
    return VB200_OK;
  }
}

namespace {
__global__ void __launch_bounds__(256) k_perm_from_keys(const unsigned long long *__restrict__ skey, int64_t n,
                                                        int *__restrict__ perm) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) perm[s] = (int)(unsigned int)skey[s];
}
}  // namespace

// Order arbitrary points (queries) the way the scene itself is laid out — by (coarse cell, fine cell, index)
// of the scene's grid, points outside clamped to the border cells — so that consecutive queries are spatial
// neighbours and the warp-cooperative search stays coherent.  d_perm[s] = index of the s-th query in that order.
int grid_order_points(const Scene *sc, const double *d_xyz, int64_t n, int *d_perm) {
    cudaStream_t st = sc->stream;
    const GridParams &g = sc->grid.p;
    const int64_t ncoarse = sc->ncoarse;
    DevBuf<int> d_ckey(st), d_ccount(st), d_cstart(st), d_cfcount(st);
    DevBuf<unsigned char> d_fbit(st);
    DevBuf<unsigned long long> d_skey(st), d_cmask(st);
    VB_CUDA(d_ckey.alloc((size_t)n));
    VB_CUDA(d_fbit.alloc((size_t)n));
    VB_CUDA(d_skey.alloc((size_t)n));
    VB_CUDA(d_ccount.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cstart.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cfcount.alloc((size_t)ncoarse + 1));
    VB_CUDA(d_cmask.alloc((size_t)ncoarse));
    VB_CUDA(cudaMemsetAsync(d_ccount.p, 0, sizeof(int) * ((size_t)ncoarse + 1), st));
    k_keys<<<div_up(n, kTpb), kTpb, 0, st>>>(g, d_xyz, n, d_ckey.p, d_fbit.p, d_ccount.p);
    VB_TRY(exclusive_scan_i32(d_ccount.p, d_cstart.p, ncoarse + 1, nullptr, st));
    VB_CUDA(cudaMemsetAsync(d_ccount.p, 0, sizeof(int) * ((size_t)ncoarse + 1), st));
    k_scatter<<<div_up(n, kTpb), kTpb, 0, st>>>(d_ckey.p, d_fbit.p, n, d_cstart.p, d_ccount.p, d_skey.p);
    k_sort_cells<<<div_up(ncoarse * 32, 256), 256, 0, st>>>((int)ncoarse, d_cstart.p, d_skey.p, d_cmask.p, d_cfcount.p);
    k_perm_from_keys<<<div_up(n, 256), 256, 0, st>>>(d_skey.p, n, d_perm);
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

void scene_free(Scene *sc) {
    if (!sc) return;
    // back to the pool, in stream order behind whatever the scene's streams still have queued
    if (sc->stream2) cudaStreamSynchronize(sc->stream2);
    const void *arrays[] = {sc->grid.coarse, sc->grid.fstart, sc->grid.hi, sc->grid.xyz, sc->grid.nrm, sc->grid.orig};
    for (const void *a : arrays)
        if (a) {
            if (sc->stream) cudaFreeAsync(const_cast<void *>(a), sc->stream); else cudaFree(const_cast<void *>(a));
        }
    if (sc->stream) cudaStreamSynchronize(sc->stream);
    if (sc->stream2) cudaStreamDestroy(sc->stream2);
    if (sc->stream) cudaStreamDestroy(sc->stream);
    delete sc;
}

}  // namespace vb

// ---------------------------------------------------------------------------------------------------
using vb::Scene;

extern "C" int vb200_scene_create(const double *xyz, const double *nrm, int64_t n, double max_radius,
                                  int device, vb200_scene_t **out) {
    if (!out) return VB200_ERR_INVALID;
    *out = nullptr;
    if (!xyz || n <= 0 || n > 0x7fffffff || !(max_radius > 0.0)) return VB200_ERR_INVALID;
    VB_TRY(vb::select_device(device));
    Scene *sc = new Scene();
    sc->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        vb::set_last_error(__FILE__, __LINE__, e);
        delete sc;
        return VB200_ERR_CUDA;
    }
    int rc = vb::scene_build(sc, xyz, nrm, n, max_radius);
    if (rc != VB200_OK) {
        vb::scene_free(sc);
        return rc;
    }
    *out = reinterpret_cast<vb200_scene_t *>(sc);
    return VB200_OK;
}

extern "C" int vb200_scene_destroy(vb200_scene_t *scene) {
    if (!scene) return VB200_OK;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    cudaSetDevice(sc->device);
    vb::scene_free(sc);
    return VB200_OK;
}

extern "C" int vb200_scene_size(const vb200_scene_t *scene, int64_t *n_points, int64_t *n_coarse_cells,
                                int64_t *n_fine_cells, double *cell_size) {
    if (!scene) return VB200_ERR_INVALID;
    const Scene *sc = reinterpret_cast<const Scene *>(scene);
    if (n_points) *n_points = sc->n;
    if (n_coarse_cells) *n_coarse_cells = sc->ncoarse;
    if (n_fine_cells) *n_fine_cells = sc->nfine;
    if (cell_size) *cell_size = sc->grid.p.cell;
    return VB200_OK;
}

extern "C" void *vb200_scene_stream(const vb200_scene_t *scene) {
    return scene ? (void *)reinterpret_cast<const Scene *>(scene)->stream : nullptr;
}

extern "C" int vb200_scene_sync(const vb200_scene_t *scene) {
    if (!scene) return VB200_ERR_INVALID;
    const Scene *sc = reinterpret_cast<const Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    VB_CUDA(cudaStreamSynchronize(sc->stream));
    return VB200_OK;
}
