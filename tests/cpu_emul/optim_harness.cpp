// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/optim_device.cuh (Adam step + soft target update in one
// launch) for the host through cuda_shim.h.
#include "cuda_shim.h"

inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

#include "optim_device.cuh"

extern "C" int emul_adam_soft(float *const *p, const float *const *g, float *const *m, float *const *v, float *const *tgt,
                              const int64_t *sizes, int n, float *step, float lr, float b1, float b2, float eps, float wd,
                              double rate) {
    AdamSoftList l{};
    l.n = n;
    l.start[0] = 0;
    for (int k = 0; k < n; ++k) {
        l.p[k] = p[k]; l.g[k] = g[k]; l.m[k] = m[k]; l.v[k] = v[k]; l.t[k] = tgt[k];
        l.start[k + 1] = l.start[k] + sizes[k];
    }
    const int64_t total = l.start[n];
    static unsigned ticket = 0;                          // zero between launches: the last CTA resets it
    shim_launch((unsigned)((total + 255) / 256 > 3 ? 3 : (total + 255) / 256), 256, [&] {      // grid-stride: fewer CTAs than work
        adam_soft_kernel(l, lr, b1, b2, eps, wd, step, &ticket, (float)(1.0 - rate), (float)rate);
    });
    return ticket == 0 ? 0 : 7;                          // (the kernel advanced *step itself)
}
