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

// ---- categorical TD target of RAINBOW (csrc/dist_device.cuh), launched with the geometry and the scalar
// conversions of mb_td_rainbow (csrc/td.cu)
#include "dist_device.cuh"

extern "C" int emul_td_rainbow(const float *q, const float *qn, const float *qtn, int64_t batch, int A, int N,
                               const float *support, const int64_t *action, const float *value, const uint8_t *terminal,
                               const double *isw, const double *scale, double v_min, double v_max, double discount_n, float *target_dist,
                               float *loss_row, float *abs_err, double *loss_elem, float *grad_q) {
    if (N < 2 || N > kDistMaxAtoms) return -1;
    DistArgs a{q, qn, qtn, support, batch, A, N, action, value, terminal, isw, scale, (float)(v_min + 1e-6), (float)(v_max - 1e-6),
               (float)v_min, (float)((v_max - v_min) / ((double)N - 1.0)), (float)discount_n, target_dist, loss_row, abs_err,
               loss_elem, grad_q};
    shim_launch((unsigned)batch, kDistThreads, [&] { td_dist_kernel(a); });
    return 0;
}
