// Status strings, CUDA error capture and per-device staging for the *_host entry points.
#include <atomic>
#include <mutex>

#include "mb_common.cuh"

namespace mb {

static thread_local char g_cuda_error[512] = "";

void set_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s (%s) at %s:%d", cudaGetErrorName(e),
             cudaGetErrorString(e), file, line);
    cudaGetLastError();  // clear the sticky-less error so later calls can proceed
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool carveout_max() {
    static const bool on = [] {
        const char *e = getenv("MB_CARVEOUT");
        return e && !strcmp(e, "max");
    }();
    return on;
}
long long launches() { return g_launches.load(std::memory_order_relaxed); }

static std::mutex g_staging_mu;
static Staging g_staging[64];

Staging *staging_for(int device) {
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_staging_mu);
    Staging *s = &g_staging[device];
    if (s->device < 0) {
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        s->device = device;
    }
    return s;
}

int Staging::ensure(size_t host_bytes, size_t dev_bytes) {
    if (host_bytes > h_bytes) {
        if (h) cudaFreeHost(h);
        h = nullptr;
        h_bytes = 0;
        size_t want = host_bytes + host_bytes / 2 + 4096;
        if (cudaMallocHost((void **)&h, want) != cudaSuccess) return MB_ERR_NOMEM;
        h_bytes = want;
    }
    if (dev_bytes > d_bytes) {
        if (d) cudaFree(d);
        d = nullptr;
        d_bytes = 0;
        size_t want = dev_bytes + dev_bytes / 2 + 4096;
        if (cudaMalloc((void **)&d, want) != cudaSuccess) return MB_ERR_NOMEM;
        d_bytes = want;
    }
    return MB_OK;
}

}  // namespace mb

extern "C" {



int mb_abi_version(void) { return MB_ABI_VERSION; }

long long mb_launch_count(void) { return mb::launches(); }

const char *mb_status_string(int status) {
    switch (status) {
        case MB_OK: return "ok";
        case MB_ERR_CUDA: return "CUDA runtime error";
        case MB_ERR_INDEX_OOB: return "Index has elements out of boundary!";
        case MB_ERR_DIM_MISMATCH: return "Dimension of weights and indexes doesn't match!";
        case MB_ERR_SIZE_MISMATCH: return "Weights size must match tree size!";
        case MB_ERR_INVALID: return "invalid argument";
        case MB_ERR_NOMEM: return "out of device or pinned memory";
        case MB_ERR_NO_DEVICE: return "no CUDA device available (machin_b200 has no CPU fallback)";
        default: return "unknown status";
    }
}

const char *mb_last_cuda_error(void) { return mb::g_cuda_error; }

int mb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // extern "C"
