// Per-row double-DQN target / priority math, shared by td.cu (stand-alone kernel) and mlp.cu (fused program).
// dqn_per.py:105-139 + dqn.py:497-501; float32, every operation rounded separately in the reference's order.
#pragma once
#include <stdint.h>

struct TdDqnArgs {
    const float *q, *qn, *qtn;
    int64_t batch;
    int A;
    const int64_t *action;
    const float *reward;
    const uint8_t *terminal;
    const double *isw;
    float discount;
    int mode;
    float *y_out, *av_out, *abs_out;
    double *loss_out;
    float *grad_q;
};

__device__ __forceinline__ void td_dqn_row(const TdDqnArgs &t, int64_t b) {
    const int A = t.A;
    // a* = argmax (first maximum, like t.max(...)[1]) over the online net ("double") or the target net
    const float *sel = (t.mode == 0 ? t.qn : t.qtn) + b * A;
    int best = 0;
    float best_v = sel[0];
    for (int a = 1; a < A; ++a) {
        const float v = sel[a];
        if (v > best_v) { best_v = v; best = a; }
    }
    const float next_value = t.qtn[b * A + best];
    // reward + discount * ~terminal * next_value  (dqn.py:501): (discount * mask) * next, then + reward
    const float mask = t.terminal[b] ? 0.0f : 1.0f;
    const float y = __fadd_rn(t.reward[b], __fmul_rn(__fmul_rn(t.discount, mask), next_value));
    if (t.y_out) t.y_out[b] = y;
    if (!t.q) return;
    int64_t act = t.action[b];
    act = act < 0 ? 0 : (act >= A ? A - 1 : act);
    const float av = t.q[b * A + act];
    const float diff = __fsub_rn(av, y);
    if (t.av_out) t.av_out[b] = av;
    if (t.abs_out) t.abs_out[b] = fabsf(diff);                          // sum(|q - y|, dim=1) with one column
    const double w = t.isw ? t.isw[b] : 1.0;
    if (t.loss_out) t.loss_out[b] = (double)__fmul_rn(diff, diff) * w;  // MSE(reduction="none") * is_weight (f64)
    if (t.grad_q) {
        // d mean_b(loss_elem) / d q[b, a]
        const float g = (float)(2.0 * (double)diff * w / (double)t.batch);
        for (int a = 0; a < A; ++a) t.grad_q[b * A + a] = (a == act) ? g : 0.0f;
    }
}
