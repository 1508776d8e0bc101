// runtime.cu — status strings, last-error text, device selection.
#include "common.cuh"

namespace vb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char *file, int line, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s:%d: %s (%s)", file, line, cudaGetErrorString(e),
             cudaGetErrorName(e));
    (void)cudaGetLastError();  // clear the sticky-free error state
}

int select_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        if (e != cudaSuccess) set_last_error(__FILE__, __LINE__, e);
        return VB200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return VB200_ERR_NO_DEVICE;
    VB_CUDA(cudaSetDevice(device));
    // keep stream-ordered allocations cached in the pool between calls (see DevBuf)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
    return VB200_OK;
}

}  // namespace vb

extern "C" int vb200_version(void) { return VB200_VERSION; }

extern "C" const char *vb200_strerror(int status) {
    switch (status) {
        case VB200_OK: return "ok";
        case VB200_ERR_INVALID: return "invalid argument";
        case VB200_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case VB200_ERR_CUDA: return "CUDA runtime error (see vb200_last_error)";
        case VB200_ERR_NOMEM: return "out of memory";
        case VB200_ERR_NORMALS: return "point-to-plane estimation requires normals on source and target";
        case VB200_ERR_DISTANCE: return "invalid max_correspondence_distance";
        default: return "unknown status";
    }
}

extern "C" const char *vb200_last_error(void) { return vb::g_last_error; }

extern "C" int vb200_release_cached_memory(int device) {
    VB_TRY(vb::select_device(device));
    cudaMemPool_t pool;
    VB_CUDA(cudaDeviceSynchronize());
    VB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    VB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return VB200_OK;
}

extern "C" int vb200_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return count;
}
