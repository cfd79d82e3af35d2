// TEST INFRASTRUCTURE ONLY -- just enough of the CUDA execution model to compile a kernel's DEVICE SOURCE with
// g++ and run it on the host: one CTA at a time, every CUDA thread an OS thread, __syncthreads() a std::barrier,
// shared memory a per-CTA buffer.  It exists so that indexing / data-flow mistakes in kernels written while no
// GPU was available can be caught in the build container.  It is not a CPU path of the library: nothing under
// machin_b200/ includes it and it is compiled only by tests/.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static      /* function-scope shared variables: one CTA runs at a time */

struct shim_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local shim_dim3 threadIdx, blockIdx;
inline shim_dim3 blockDim, gridDim;
inline std::barrier<> *shim_cta_barrier = nullptr;

inline void __syncthreads() { shim_cta_barrier->arrive_and_wait(); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }

struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
inline float2 __ldg(const float2 *p) { return *p; }
inline float __ldg(const float *p) { return *p; }
inline float4 __ldg(const float4 *p) { return *p; }
inline float __ldcg(const float *p) { return *p; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }   // volatile: no contraction
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }

// Run `kernel()` for every thread of every CTA of a 1-D grid, CTAs one after the other.
inline void shim_launch(unsigned grid, unsigned block, const std::function<void()> &kernel) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> bar((std::ptrdiff_t)block);
        shim_cta_barrier = &bar;
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel();
            });
        for (auto &th : threads) th.join();
    }
    shim_cta_barrier = nullptr;
}
