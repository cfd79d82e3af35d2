// Categorical (distributional) TD target of RAINBOW for one batch row per CTA: greedy next action by expectation,
// Bellman update of every atom, projection onto the fixed support, cross-entropy row loss, priority and
// d loss / d q_dist -- rainbow.py:212-306 + dqn.py:497-501 in ONE pass instead of ~30 eager CPU tensor ops.
//
// float32, every operation rounded separately in the reference's order.  The projection adds the lower-neighbour
// contributions in atom order and then the upper-neighbour ones, exactly like the two sequential index_add_ calls
// (:279-280), so target_dist is bit-identical to CPU torch given the same network outputs -- including the rows
// whose target sits exactly on an atom (floor == ceil: both weights are 0 and the mass is dropped, :257-262).
#pragma once
#include <stdint.h>

constexpr int kDistThreads = 128;
constexpr int kDistMaxAtoms = 256;

struct DistArgs {
    const float *q, *qn, *qtn;   // [B, A, N]: online net on s, online net on s', target net on s' (soft-maxed)
    const float *support;        // [N] = t.linspace(v_min, v_max, N), float32
    int64_t batch;
    int A, N;
    const int64_t *action;       // [B]
    const float *value;          // [B] n-step value stored with the transition (:186-196)
    const uint8_t *terminal;     // [B]
    const double *isw;           // [B] or null (all ones)
    const double *scale;         // null, or 1 value = sum_i is_weight_i / B^2 computed by the caller (large batches:
                                 // otherwise EVERY row's CTA walks the whole is_weight vector)
    float lo, hi;                // clamp bounds: (float)(v_min + 1e-6), (float)(v_max - 1e-6)
    float vmin, delta_z;         // (float)v_min, (float)((v_max - v_min) / (N - 1.0))
    float gamma_n;               // (float)(discount ** reward_future_steps)
    float *target_dist;          // [B, N] or null
    float *loss_row;             // [B] or null
    float *abs_out;              // [B] = |row loss| + 1e-6 (the new priority)
    double *loss_out;            // [B] or null; mean over the batch = the reference's scalar loss
    float *grad_q;               // [B, A, N] or null
};

__global__ void __launch_bounds__(kDistThreads) td_dist_kernel(DistArgs t) {
    __shared__ float s_lw[kDistMaxAtoms], s_uw[kDistMaxAtoms], s_tgt[kDistMaxAtoms], s_term[kDistMaxAtoms];
    __shared__ int s_l[kDistMaxAtoms], s_u[kDistMaxAtoms];
    __shared__ float s_best_v[kDistThreads];
    __shared__ int s_best_a[kDistThreads];
    __shared__ double s_part[kDistThreads];
    __shared__ int s_next;
    __shared__ double s_scale;
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x, A = t.A, N = t.N;

    // greedy action of the ONLINE net on s' (act_discrete, :110-126): expectation over the support, first maximum
    float best_v = 0.0f;
    int best_a = -1;
    for (int a = tid; a < A; a += kDistThreads) {
        const float *p = t.qn + (b * A + a) * N;
        float e = 0.0f;
        for (int k = 0; k < N; ++k) e = __fadd_rn(e, __fmul_rn(t.support[k], p[k]));
        if (best_a < 0 || e > best_v) { best_v = e; best_a = a; }
    }
    s_best_v[tid] = best_v;
    s_best_a[tid] = best_a;
    // sum_i is_weight_i / B^2: what the mean over the [B, B] product (:303) sends back to every row
    const double inv = 1.0 / ((double)t.batch * (double)t.batch);
    double part = 0.0;
    if (t.scale) {
        if (tid == 0) part = *t.scale;
    } else if (!t.isw) {
        if (tid == 0) part = (double)t.batch * inv;
    } else {
        for (int64_t i = tid; i < t.batch; i += kDistThreads) part += t.isw[i] * inv;
    }
    s_part[tid] = part;
    __syncthreads();
    if (tid == 0) {
        float v = s_best_v[0];
        int a = s_best_a[0];
        double total = s_part[0];
        for (int i = 1; i < kDistThreads; ++i) {
            total += s_part[i];
            if (s_best_a[i] < 0) continue;
            if (a < 0 || s_best_v[i] > v || (s_best_v[i] == v && s_best_a[i] < a)) { v = s_best_v[i]; a = s_best_a[i]; }
        }
        s_next = a < 0 ? 0 : a;
        s_scale = total;
    }
    __syncthreads();

    // T_z = value + (discount^n * ~terminal) * z_j, clamped; b_j = (T_z - v_min) / delta_z; neighbours and weights
    const float cont = __fmul_rn(t.gamma_n, t.terminal[b] ? 0.0f : 1.0f);
    const float value = t.value[b];
    const float *p_next = t.qtn + (b * A + s_next) * N;
    for (int j = tid; j < N; j += kDistThreads) {
        float tz = __fadd_rn(value, __fmul_rn(cont, t.support[j]));
        tz = fminf(fmaxf(tz, t.lo), t.hi);
        const float bj = __fdiv_rn(__fsub_rn(tz, t.vmin), t.delta_z);
        const float lower = floorf(bj), upper = ceilf(bj);
        const float p = p_next[j];
        s_l[j] = (int)lower;
        s_u[j] = (int)upper;
        s_lw[j] = __fmul_rn(__fsub_rn(upper, bj), p);     // l_weight = u_dist * p
        s_uw[j] = __fmul_rn(__fsub_rn(bj, lower), p);     // u_weight = l_dist * p
    }
    __syncthreads();

    int64_t act = t.action[b];
    act = act < 0 ? 0 : (act >= A ? A - 1 : act);
    const float *q_sel = t.q + (b * A + act) * N;
    for (int k = tid; k < N; k += kDistThreads) {
        float m = 0.0f;
        for (int j = 0; j < N; ++j)
            if (s_l[j] == k) m = __fadd_rn(m, s_lw[j]);
        for (int j = 0; j < N; ++j)
            if (s_u[j] == k) m = __fadd_rn(m, s_uw[j]);
        s_tgt[k] = m;
        if (t.target_dist) t.target_dist[b * N + k] = m;
        s_term[k] = __fmul_rn(m, logf(__fadd_rn(q_sel[k], 1e-6f)));     // target * (q + 1e-6).log()  (:291)
    }
    __syncthreads();

    if (tid == 0) {
        float acc = 0.0f;
        for (int k = 0; k < N; ++k) acc = __fadd_rn(acc, s_term[k]);
        const float row = -acc;
        if (t.loss_row) t.loss_row[b] = row;
        if (t.abs_out) t.abs_out[b] = __fadd_rn(fabsf(row), 1e-6f);    // :295
        if (t.loss_out) t.loss_out[b] = (double)row * s_scale * (double)t.batch;
    }
    if (t.grad_q) {
        const float g = -(float)s_scale;
        float *gq = t.grad_q + b * A * N;
        for (int i = tid; i < A * N; i += kDistThreads) {
            const int a = i / N, k = i - a * N;
            gq[i] = a == act ? __fdiv_rn(__fmul_rn(g, s_tgt[k]), __fadd_rn(q_sel[k], 1e-6f)) : 0.0f;
        }
    }
}
