// runtime.cu — status strings, last-error text, device selection.
#include <mutex>
#include <thread>

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

// ---- staged host-to-device copies -----------------------------------------------------------------------------
namespace {
constexpr size_t kStageChunk = 4u << 20;     // bytes per staging slot (1-4 MB x 4-8 workers all measure the same)
constexpr int kStageWorkers = 4;             // host threads (and slots x 2) per copy
constexpr size_t kStageMinBytes = 16u << 20; // below this the driver's own staging is as good
struct StagePool {                           // one per process, per device; built on first use, kept
    std::mutex mu;
    int device = -1;
    void *slot[2 * kStageWorkers] = {};
    cudaEvent_t freed[2 * kStageWorkers] = {};
    cudaStream_t copy[kStageWorkers] = {};
    cudaEvent_t start = nullptr, done[kStageWorkers] = {};
    bool ok = false;
    bool init(int dev) {
        if (ok && device == dev) return true;
        if (ok) return false;  // (a second device takes the plain path)
        device = dev;
        for (int i = 0; i < 2 * kStageWorkers; i++) {
            if (cudaHostAlloc(&slot[i], kStageChunk, cudaHostAllocDefault) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        for (int i = 0; i < kStageWorkers; i++) {
            if (cudaStreamCreateWithFlags(&copy[i], cudaStreamNonBlocking) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaEventCreateWithFlags(&start, cudaEventDisableTiming) != cudaSuccess) return false;
        ok = true;
        return true;
    }
};
StagePool g_stage;
}  // namespace

cudaError_t h2d_async(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    bool staged = bytes >= kStageMinBytes;
    if (staged) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, h_src) != cudaSuccess) { (void)cudaGetLastError(); staged = false; }
        else staged = at.type == cudaMemoryTypeUnregistered;  // plain pageable host memory
    }
    int dev = 0;
    if (staged && cudaGetDevice(&dev) != cudaSuccess) staged = false;
    if (!staged) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
    std::unique_lock<std::mutex> lock(g_stage.mu);
    if (!g_stage.init(dev)) {
        (void)cudaGetLastError();
        lock.unlock();
        return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
    }
    // the transfers start after everything already queued on `st` (the destination may still be in use there) ...
    cudaError_t e = cudaEventRecord(g_stage.start, st);
    if (e != cudaSuccess) return e;
    const size_t nchunk = (bytes + kStageChunk - 1) / kStageChunk;
    cudaError_t werr[kStageWorkers];
    std::thread th[kStageWorkers];
    for (int w = 0; w < kStageWorkers; w++) {
        werr[w] = cudaSuccess;
        th[w] = std::thread([&, w]() {
            cudaError_t r = cudaSetDevice(dev);
            if (r == cudaSuccess) r = cudaStreamWaitEvent(g_stage.copy[w], g_stage.start, 0);
            int k = 0;  // this worker's two slots alternate
            for (size_t c = (size_t)w; c < nchunk && r == cudaSuccess; c += kStageWorkers, k ^= 1) {
                const int s = 2 * w + k;
                const size_t off = c * kStageChunk, len = bytes - off < kStageChunk ? bytes - off : kStageChunk;
                r = cudaEventSynchronize(g_stage.freed[s]);  // the slot's previous transfer has left it
                if (r != cudaSuccess) break;
                memcpy(g_stage.slot[s], (const char *)h_src + off, len);
                r = cudaMemcpyAsync((char *)d_dst + off, g_stage.slot[s], len, cudaMemcpyHostToDevice, g_stage.copy[w]);
                if (r == cudaSuccess) r = cudaEventRecord(g_stage.freed[s], g_stage.copy[w]);
            }
            if (r == cudaSuccess) r = cudaEventRecord(g_stage.done[w], g_stage.copy[w]);
            werr[w] = r;
        });
    }
    for (int w = 0; w < kStageWorkers; w++) th[w].join();
    // ... and `st` continues once every worker's last chunk has landed
    for (int w = 0; w < kStageWorkers; w++) {
        if (werr[w] != cudaSuccess) return werr[w];
        e = cudaStreamWaitEvent(st, g_stage.done[w], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
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
