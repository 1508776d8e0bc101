// scan.cu — device-wide exclusive prefix sum (reduce-then-scan, 3 launches) used by the grid build,
// the source-cloud bucket sort and the voxel hash.  Plain CUDA; HBM-bound: reads the input twice,
// writes it once (12 B/element).
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace vb {

namespace {

constexpr int kTpb = 512;
constexpr int kIpt = 8;
constexpr int kTile = kTpb * kIpt;

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// inclusive block scan of one value per thread; returns inclusive value, *total = block sum
__device__ __forceinline__ int block_incl_scan(int v, int *total) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? wsum[lane] : 0;
        w = warp_incl_scan(w);
        wsum[lane] = w;
    }
    __syncthreads();
    int off = warp ? wsum[warp - 1] : 0;
    *total = wsum[nw - 1];
    __syncthreads();
    return inc + off;
}

__global__ void __launch_bounds__(kTpb) k_tile_sums(const int *__restrict__ in, int64_t n,
                                                    int *__restrict__ tsum) {
    int64_t base = (int64_t)blockIdx.x * kTile;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kIpt; k++) {
        int64_t i = base + (int64_t)k * kTpb + threadIdx.x;
        if (i < n) s += in[i];
    }
    int tot;
    block_incl_scan(s, &tot);
    if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tile_sums(int *tsum, int nt, int *total) {
    int carry = 0;
    for (int base = 0; base < nt; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nt ? tsum[i] : 0;
        int tot;
        int inc = block_incl_scan(v, &tot);
        if (i < nt) tsum[i] = carry + inc - v;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(kTpb) k_apply(const int *__restrict__ in, int *__restrict__ out,
                                                int64_t n, const int *__restrict__ tsum) {
    // thread t owns kIpt consecutive items so the result is an exclusive scan in index order
    int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kIpt;
    int v[kIpt];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kIpt; k++) {
        int64_t i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    int tot;
    int inc = block_incl_scan(s, &tot);
    int run = tsum[blockIdx.x] + inc - s;
#pragma unroll
    for (int k = 0; k < kIpt; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
}

constexpr int kBoxTpb = 256;
__global__ void __launch_bounds__(kBoxTpb) k_bbox(const double *__restrict__ xyz, int64_t n,
                                               double *__restrict__ part /* [grid][6] */) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double v = xyz[3 * i + a];
            lo[a] = fmin(lo[a], v);
            hi[a] = fmax(hi[a], v);
        }
    }
    __shared__ double s[kBoxTpb / 32][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++)
        for (int o = 16; o; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if (lane == 0)
        for (int a = 0; a < 3; a++) { s[warp][a] = lo[a]; s[warp][3 + a] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[0][threadIdx.x];
        for (int w = 1; w < kBoxTpb / 32; w++)
            v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

}  // namespace

int exclusive_scan_i32(const int *d_in, int *d_out, int64_t n, int *d_total, cudaStream_t stream) {
    if (n <= 0) {
        if (d_total) VB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), stream));
        return VB200_OK;
    }
    int nt = div_up(n, kTile);
    int *tsum = nullptr;
    VB_CUDA(cudaMallocAsync((void **)&tsum, sizeof(int) * (size_t)nt, stream));
    k_tile_sums<<<nt, kTpb, 0, stream>>>(d_in, n, tsum);
    k_scan_tile_sums<<<1, 1024, 0, stream>>>(tsum, nt, d_total);
    k_apply<<<nt, kTpb, 0, stream>>>(d_in, d_out, n, tsum);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaFreeAsync(tsum, stream));
    return VB200_OK;
}

int device_bbox(const double *d_xyz, int64_t n, double lo[3], double hi[3], cudaStream_t stream) {
    const int nb = std::min(div_up(n, kBoxTpb), 4 * kNumSMsB200);
    DevBuf<double> d_part(stream);
    VB_CUDA(d_part.alloc(6 * (size_t)nb));
    k_bbox<<<nb, kBoxTpb, 0, stream>>>(d_xyz, n, d_part.p);
    VB_CUDA(cudaGetLastError());
    std::vector<double> part(6 * (size_t)nb);
    VB_CUDA(cudaMemcpyAsync(part.data(), d_part.p, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));
    for (int a = 0; a < 3; a++) { lo[a] = 1e300; hi[a] = -1e300; }
    for (int b = 0; b < nb; b++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], part[6 * b + a]);
            hi[a] = std::max(hi[a], part[6 * b + 3 + a]);
        }
    return VB200_OK;
}

}  // namespace vb
