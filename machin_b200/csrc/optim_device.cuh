// Adam step + soft target update in one launch: device code only (no CUDA runtime calls).  Included by mlp.cu
// inside its anonymous namespace and by tests/cpu_emul/optim_harness.cpp (g++ against the CUDA shim).
//
// torch.optim.Adam (amsgrad off, maximize off) element by element exactly like adam_kernel in mlp.cu, then -- the
// updated parameter still in a register -- utils.py:8-27's target = target * (1 - rate) + param * rate with the
// two scalars rounded to float32 first, like soft_update_multi_kernel in td.cu.  Saves the second pass over the
// parameters and one launch per learner update.  The step counter is advanced by the kernel itself: every CTA reads
// it first, and the CTA that finishes LAST (an atomic ticket) writes t + 1 -- no CTA can still need the old value
// then -- so the separate one-thread increment launch of mb_adam_step_multi is not needed either.
#pragma once
#include <stdint.h>

#include "machin_b200.h"

struct AdamSoftList {
    float *p[MB_MAX_SOFT_TENSORS];
    const float *g[MB_MAX_SOFT_TENSORS];
    float *m[MB_MAX_SOFT_TENSORS];
    float *v[MB_MAX_SOFT_TENSORS];
    float *t[MB_MAX_SOFT_TENSORS];          // target-network tensor matching p[k] (never NULL)
    int64_t start[MB_MAX_SOFT_TENSORS + 1];
    int n;
};

__global__ void __launch_bounds__(256) adam_soft_kernel(AdamSoftList l, float lr, float beta1, float beta2, float eps,
                                                        float weight_decay, float *step_count, unsigned *done, float keep,
                                                        float rate) {
    const int64_t total = l.start[l.n];
    const float t = *step_count + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        while (i >= l.start[k + 1]) ++k;
        const int64_t j = i - l.start[k];
        float grad = l.g[k][j];
        const float p = l.p[k][j];
        if (weight_decay != 0.0f) grad = fmaf(p, weight_decay, grad);
        const float m = __fadd_rn(l.m[k][j], __fmul_rn(__fsub_rn(grad, l.m[k][j]), __fsub_rn(1.0f, beta1)));   // lerp, like torch
        const float v = __fadd_rn(__fmul_rn(l.v[k][j], beta2), __fmul_rn(__fmul_rn(__fsub_rn(1.0f, beta2), grad), grad));
        l.m[k][j] = m;
        l.v[k][j] = v;
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
        const float p_new = __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(m, denom)));
        l.p[k][j] = p_new;
        l.t[k][j] = __fadd_rn(__fmul_rn(l.t[k][j], keep), __fmul_rn(p_new, rate));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {     // every CTA has read *step_count by now
            *done = 0;
            *step_count = t;
        }
    }
}
