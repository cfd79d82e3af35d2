// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/rows_program.cuh (the device source of dqn_rows_kernel)
// for the host through cuda_shim.h and exposes one C entry point for tests/test_rows_kernel_emulated_cpu.py.
#include "cuda_shim.h"

// The kernel declares its dynamic shared memory as `extern __shared__ __align__(16) float rsm[];`: with
// __shared__ spelled as nothing that is a block-scope declaration of this array (one CTA runs at a time).
__attribute__((aligned(16))) float rsm[64 * 1024];
#undef __shared__
#define __shared__
#include "rows_program.cuh"

extern "C" int emul_dqn_rows(int L, const int *sizes, int batch, const float *const *W, const float *const *Wt,
                             const float *const *tWt, const float *const *b, const float *const *tb,
                             const float *state, const float *next_state, float *const *h, float *qn, float *qtn,
                             float *const *dz, const int64_t *action, const float *reward, const uint8_t *terminal,
                             const double *isw, float discount, int mode, float *y, float *av, float *abs_err,
                             double *loss, int cpt) {
    if (L < 1 || L > RL) return -1;
    RowsArgs a{};
    a.L = L;
    a.batch = batch;
    for (int l = 0; l <= L; ++l) a.sizes[l] = sizes[l];
    for (int l = 0; l < L; ++l) {
        a.W[l] = W[l]; a.Wt[l] = Wt[l]; a.tWt[l] = tWt[l]; a.b[l] = b[l]; a.tb[l] = tb[l];
        a.h[l] = h[l]; a.dz[l] = dz[l];
    }
    a.state = state;
    a.next_state = next_state;
    a.qn = qn;
    a.qtn = qtn;
    a.td = TdDqnArgs{h[L - 1], qn, qtn, batch, sizes[L], action, reward, terminal, isw, discount, mode, y, av, abs_err,
                     loss, dz[L - 1]};
    const int xp = ((sizes[0] + 3) & ~3) + 4;
    if ((size_t)(2 * RB) * (size_t)(xp + 2 * RP) > sizeof(rsm) / sizeof(float)) return -2;
    if (cpt == 2) shim_launch((unsigned)((batch + RB - 1) / RB), RT, [&] { dqn_rows_kernel<2>(a); });
    else shim_launch((unsigned)((batch + RB - 1) / RB), RT, [&] { dqn_rows_kernel<4>(a); });
    return 0;
}
