// Shared helpers for libmachin_b200 (sm_100a).  Internal; the public surface is include/machin_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "machin_b200.h"

namespace mb {

void set_cuda_error(cudaError_t e, const char *file, int line);
void count_launch(int n);  // kernels launched by this library (mb_launch_count)
// MB_CARVEOUT=max (experiment): every kernel of the learner update asks for the largest shared-memory
// carve-out, so that consecutive kernels of one update never make an SM switch its L1 / shared-memory split.  A
// preference only: results do not depend on it.
bool carveout_max();
#define MB_PREFER_MAX_SMEM(...)                                                                              \
    do {                                                                                                     \
        static const bool mb_once_ = mb::carveout_max() &&                                                   \
            cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributePreferredSharedMemoryCarveout,                \
                                 (int)cudaSharedmemCarveoutMaxShared) == cudaSuccess;                        \
        (void)mb_once_;                                                                                      \
    } while (0)

#define MB_CUDA(expr)                                        \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            mb::set_cuda_error(_e, __FILE__, __LINE__);      \
            return MB_ERR_CUDA;                              \
        }                                                    \
    } while (0)

#define MB_LAUNCH_CHECK() MB_CUDA(cudaGetLastError())
// record n kernel launches, then check the launch status
#define MB_LAUNCHED(n)       \
    do {                     \
        mb::count_launch(n); \
        MB_LAUNCH_CHECK();   \
    } while (0)

// the CTA's dynamic shared memory as a byte array (a macro so that the host-side emulation used by the tests can
// substitute its own buffer; under nvcc this is the usual declaration)
#define MB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Order-preserving map double -> uint64 so that atomicMax works for any sign.
__host__ __device__ inline unsigned long long f64_to_key(double x) {
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b;
    memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ inline double key_to_f64(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Pinned + device staging shared by the *_host entry points of one device.
// `mu` is held for the whole of a *_host call (ensure -> memcpy -> launch -> synchronise -> copy-out): the library is
// loaded with ctypes.CDLL (GIL released), so two Python threads may enter host entry points of one GPU at once.
struct Staging {
    std::mutex mu;
    int device = -1;
    cudaStream_t stream = nullptr;
    char *h = nullptr;
    size_t h_bytes = 0;
    char *d = nullptr;
    size_t d_bytes = 0;
    int ensure(size_t host_bytes, size_t dev_bytes);
};
Staging *staging_for(int device);

}  // namespace mb
