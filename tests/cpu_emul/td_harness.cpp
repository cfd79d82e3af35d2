// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/td_device.cuh (the per-row double-DQN target / priority math
// shared by td_dqn_kernel and the fused programs) for the host through cuda_shim.h.
#include "cuda_shim.h"

#include "td_device.cuh"

extern "C" int emul_td_dqn(const float *q, const float *qn, const float *qtn, int64_t batch, int A, const int64_t *action,
                           const float *reward, const uint8_t *terminal, const double *isw, float discount, int mode,
                           float *y, float *av, float *abs_err, double *loss, float *grad_q) {
    TdDqnArgs a{q, qn, qtn, batch, A, action, reward, terminal, isw, discount, mode, y, av, abs_err, loss, grad_q};
    shim_launch((unsigned)((batch + 255) / 256), 256, [&] {      // the geometry of td_dqn_kernel (csrc/td.cu)
        const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (b < a.batch) td_dqn_row(a, b);
    });
    return 0;
}
