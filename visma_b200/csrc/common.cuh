// common.cuh — shared host/device helpers for libvisma_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/visma_b200.h"

namespace vb {

// ---- error plumbing: nothing throws across the C ABI -------------------------------------------------
void set_last_error(const char *file, int line, cudaError_t e);

#define VB_CUDA(expr)                                        \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            vb::set_last_error(__FILE__, __LINE__, _e);      \
            return VB200_ERR_CUDA;                           \
        }                                                    \
    } while (0)

#define VB_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != VB200_OK) return _s; \
    } while (0)

int select_device(int device);  // validates + cudaSetDevice; VB200_ERR_NO_DEVICE if absent

// Host-to-device copy of a caller's buffer, ordered on `st` like cudaMemcpyAsync.  A pinned source goes straight to the
// copy engine.  A large PAGEABLE source — what a std::vector<Eigen::Vector3d> is — would be staged by the driver through
// its own bounce buffer by one thread (~10 GB/s: 9 ms for a 2 M-point scene with normals); here a few host threads copy
// it chunk by chunk into pinned staging slots and queue each chunk's transfer as soon as it is staged (runtime.cu).
cudaError_t h2d_async(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);

constexpr int kNumSMsB200 = 148;

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device buffer with RAII so early returns do not leak.  Constructed with a stream it uses the
// stream-ordered allocator (cudaMallocAsync / cudaFreeAsync): with the pool's release threshold raised in
// select_device() repeated calls reuse the same memory instead of paying cudaMalloc/cudaFree (which
// synchronise the device) on every vb200_icp_run.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t st = nullptr;
    bool async = false;
    DevBuf() {}
    explicit DevBuf(cudaStream_t s) : st(s), async(true) {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return async ? cudaMallocAsync((void **)&p, count * sizeof(T), st) : cudaMalloc((void **)&p, count * sizeof(T));
    }
    void release() {
        if (p) {
            if (async) cudaFreeAsync(p, st); else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    T *take() {
        T *q = p;
        p = nullptr;
        n = 0;
        return q;
    }
};

// exclusive prefix sum of int32 counts on `stream`; out may alias in.  n up to 2^31-1.
// d_total (nullable) receives the grand total.  Implemented in scan.cu.
int exclusive_scan_i32(const int *d_in, int *d_out, int64_t n, int *d_total, cudaStream_t stream);

// axis-aligned bounding box of n double3 points resident on the device (synchronises `stream`)
int device_bbox(const double *d_xyz, int64_t n, double lo[3], double hi[3], cudaStream_t stream);

}  // namespace vb
