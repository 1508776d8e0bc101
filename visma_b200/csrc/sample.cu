// sample.cu — vb200_sample_mesh: replaces feh::SamplePointCloudFromMesh (include/geometry.h:29-64), the
// O(samples x faces) linear scan that builds every ICP source cloud (src/evaluation.cpp:250-256,
// src/annotation.cpp:126; 50 000 x 4 999 steps per object in the reference).
//
// Area-weighted face choice by binary search in the cumulative-area table + a uniform point in the triangle,
// one thread per sample, counter-based Philox4x32-10 random numbers keyed by (seed, sample index) so the cloud
// is reproducible and independent of the launch shape.  Deliberate deviations from the reference, which cannot
// be matched sample-for-sample anyway because it seeds from the wall clock (geometry.h:47):
//   - its scan tests area[k] <= r < area[k+1] and then uses face k (off by one, and r < area[0] yields no
//     sample at all, geometry.h:52-57): here the face whose interval contains r is used and exactly
//     n samples are produced;
//   - it places the sample at v0 + a (v1-v0) + b (v2-v0) with independent a, b in [0,1) — a parallelogram,
//     half of which lies outside the triangle (geometry.h:54-56): here (a, b) with a + b > 1 is reflected
//     back into the triangle.
// Parity with the reference is therefore statistical (face histogram ~ area, samples on their triangles).
#include <math.h>

#include <vector>

#include "common.cuh"

namespace vb {

namespace {

__host__ __device__ inline void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned out[4]) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256) k_face_area(const float *__restrict__ V, const int *__restrict__ F, int nF,
                                                   int nV, double *__restrict__ area) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const int i0 = F[3 * f], i1 = F[3 * f + 1], i2 = F[3 * f + 2];
    double a = 0.0;
    if ((unsigned)i0 < (unsigned)nV && (unsigned)i1 < (unsigned)nV && (unsigned)i2 < (unsigned)nV) {
        const double e1[3] = {(double)V[3 * i1] - V[3 * i0], (double)V[3 * i1 + 1] - V[3 * i0 + 1], (double)V[3 * i1 + 2] - V[3 * i0 + 2]};
        const double e2[3] = {(double)V[3 * i2] - V[3 * i0], (double)V[3 * i2 + 1] - V[3 * i0 + 1], (double)V[3 * i2 + 2] - V[3 * i0 + 2]};
        const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        a = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);  // geometry.h:38
    }
    area[f] = a;
}

__global__ void __launch_bounds__(256) k_sample(const float *__restrict__ V, const int *__restrict__ F, int nF,
                                                const double *__restrict__ cdf, int64_t n, unsigned long long seed,
                                                double *__restrict__ out_xyz, double *__restrict__ out_nrm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned rnd[4];
    philox4x32_10((unsigned)i, (unsigned)(i >> 32), 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32), rnd);
    const double r = ((double)rnd[0] * 4294967296.0 + (double)rnd[1]) * (1.0 / 18446744073709551616.0) * cdf[nF - 1];
    int lo = 0, hi = nF - 1;  // first face whose cumulative area exceeds r
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > r) hi = mid; else lo = mid + 1;
    }
    const int f = lo;
    double a = (double)rnd[2] * (1.0 / 4294967296.0), b = (double)rnd[3] * (1.0 / 4294967296.0);
    if (a + b > 1.0) { a = 1.0 - a; b = 1.0 - b; }
    const int i0 = F[3 * f], i1 = F[3 * f + 1], i2 = F[3 * f + 2];
    const double v0[3] = {V[3 * i0], V[3 * i0 + 1], V[3 * i0 + 2]};
    const double e1[3] = {(double)V[3 * i1] - v0[0], (double)V[3 * i1 + 1] - v0[1], (double)V[3 * i1 + 2] - v0[2]};
    const double e2[3] = {(double)V[3 * i2] - v0[0], (double)V[3 * i2 + 1] - v0[1], (double)V[3 * i2 + 2] - v0[2]};
    for (int k = 0; k < 3; k++) out_xyz[3 * i + k] = __dadd_rn(__dadd_rn(v0[k], __dmul_rn(a, e1[k])), __dmul_rn(b, e2[k]));
    if (out_nrm) {
        const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        const double l = sqrt(cx * cx + cy * cy + cz * cz);
        const double inv = l > 0.0 ? 1.0 / l : 0.0;
        out_nrm[3 * i] = cx * inv; out_nrm[3 * i + 1] = cy * inv; out_nrm[3 * i + 2] = cz * inv;
    }
}

}  // namespace

}  // namespace vb

extern "C" int vb200_sample_mesh(const float *V, int64_t nV, const int32_t *F, int64_t nF, int64_t n_samples,
                                 uint64_t seed, int device, double *out_xyz, double *out_nrm) {
    using namespace vb;
    if (n_samples < 0 || nV < 0 || nF < 0 || nV > 0x7fffffff || nF > 0x7fffffff) return VB200_ERR_INVALID;
    if (n_samples == 0) return VB200_OK;
    if (!V || !F || !out_xyz || nF == 0 || nV == 0) return VB200_ERR_INVALID;
    VB_TRY(select_device(device));
    cudaStream_t st;
    VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    DevBuf<float> d_V(st);
    DevBuf<int> d_F(st);
    DevBuf<double> d_area(st), d_out(st), d_nrm(st);
    VB_CUDA(d_V.alloc(3 * (size_t)nV));
    VB_CUDA(d_F.alloc(3 * (size_t)nF));
    VB_CUDA(d_area.alloc((size_t)nF));
    VB_CUDA(d_out.alloc(3 * (size_t)n_samples));
    if (out_nrm) VB_CUDA(d_nrm.alloc(3 * (size_t)n_samples));
    VB_CUDA(cudaMemcpyAsync(d_V.p, V, sizeof(float) * 3 * (size_t)nV, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_F.p, F, sizeof(int) * 3 * (size_t)nF, cudaMemcpyHostToDevice, st));
    k_face_area<<<div_up(nF, 256), 256, 0, st>>>(d_V.p, d_F.p, (int)nF, (int)nV, d_area.p);
    VB_CUDA(cudaGetLastError());
    // cumulative area table: a sequential double sum like the reference's (geometry.h:41-44), on the host —
    // F*8 bytes each way, once per mesh
    std::vector<double> cdf((size_t)nF);
    VB_CUDA(cudaMemcpyAsync(cdf.data(), d_area.p, sizeof(double) * (size_t)nF, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    double run = 0.0;
    for (int64_t f = 0; f < nF; f++) { run += cdf[f]; cdf[f] = run; }
    if (!(run > 0.0) || !std::isfinite(run)) return VB200_ERR_INVALID;  // no area to sample from
    VB_CUDA(cudaMemcpyAsync(d_area.p, cdf.data(), sizeof(double) * (size_t)nF, cudaMemcpyHostToDevice, st));
    k_sample<<<div_up(n_samples, 256), 256, 0, st>>>(d_V.p, d_F.p, (int)nF, d_area.p, n_samples, seed, d_out.p,
                                                     out_nrm ? d_nrm.p : nullptr);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(out_xyz, d_out.p, sizeof(double) * 3 * (size_t)n_samples, cudaMemcpyDeviceToHost, st));
    if (out_nrm) VB_CUDA(cudaMemcpyAsync(out_nrm, d_nrm.p, sizeof(double) * 3 * (size_t)n_samples, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    return VB200_OK;
}
