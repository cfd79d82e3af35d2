// TD-target / priority math for sm_100a: one fused pass instead of ~10 eager [B,1] ops.
//
// Replaces the arithmetic of DQNPer.update (dqn_per.py:105-139 + dqn.py:497-501), DDPGPer.update
// (ddpg_per.py:117-160 + ddpg.py:506-510) and soft_update (algorithms/utils.py:8-27).  float32 with
// each operation rounded separately (no FMA contraction) in the reference's order, so y, the
// gathered action value and abs_error are bit-identical to CPU torch given the same Q values.
#include "mb_common.cuh"
#include "td_device.cuh"
#include "dist_device.cuh"

namespace {

__global__ void __launch_bounds__(256) td_dqn_kernel(TdDqnArgs t) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < t.batch) td_dqn_row(t, b);
}

__global__ void __launch_bounds__(256) td_target_kernel(const float *__restrict__ next_value,
                                                        const float *__restrict__ reward,
                                                        const uint8_t *__restrict__ terminal, int64_t batch,
                                                        float discount, const float *__restrict__ value,
                                                        float *__restrict__ y_out, float *__restrict__ abs_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const float mask = terminal[b] ? 0.0f : 1.0f;
    const float y = __fadd_rn(reward[b], __fmul_rn(__fmul_rn(discount, mask), next_value[b]));
    if (y_out) y_out[b] = y;
    if (value && abs_out) abs_out[b] = fabsf(__fsub_rn(value[b], y));
}

// target * (1 - rate) + source * rate, evaluated like torch: scalars rounded to float32 first
__global__ void __launch_bounds__(256) soft_update_kernel(float *__restrict__ target, const float *__restrict__ source,
                                                          int64_t n, float keep, float rate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) target[i] = __fadd_rn(__fmul_rn(target[i], keep), __fmul_rn(source[i], rate));
}

struct SoftList {
    float *target[MB_MAX_SOFT_TENSORS];
    const float *source[MB_MAX_SOFT_TENSORS];
    int64_t start[MB_MAX_SOFT_TENSORS + 1];  // prefix sums of element counts
    int n;
};

// All parameter blocks of a network in one launch (one launch per tensor is launch-latency bound).
__global__ void __launch_bounds__(256) soft_update_multi_kernel(SoftList l, float keep, float rate) {
    const int64_t total = l.start[l.n];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        while (i >= l.start[k + 1]) ++k;
        const int64_t j = i - l.start[k];
        l.target[k][j] = __fadd_rn(__fmul_rn(l.target[k][j], keep), __fmul_rn(l.source[k][j], rate));
    }
}

}  // namespace

extern "C" {

int mb_td_dqn(const float *d_q, const float *d_q_next, const float *d_q_tgt_next, int64_t batch, int n_actions,
              const int64_t *d_action, const float *d_reward, const uint8_t *d_terminal,
              const double *d_is_weight, float discount, int mode, float *d_y, float *d_action_value,
              float *d_abs_error, double *d_loss_elem, float *d_grad_q, void *stream) {
    if (batch < 0 || n_actions < 1 || !d_q_tgt_next || !d_reward || !d_terminal) return MB_ERR_INVALID;
    if (mode == 0 && !d_q_next) return MB_ERR_INVALID;
    if (d_q && !d_action) return MB_ERR_INVALID;
    if (batch == 0) return MB_OK;
    TdDqnArgs a{d_q, d_q_next, d_q_tgt_next, batch, n_actions, d_action, d_reward, d_terminal, d_is_weight,
                discount, mode, d_y, d_action_value, d_abs_error, d_loss_elem, d_grad_q};
    MB_PREFER_MAX_SMEM(td_dqn_kernel);
    td_dqn_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_td_rainbow(const float *d_q_dist, const float *d_next_q_dist, const float *d_target_next_q_dist, int64_t batch,
                  int n_actions, int n_atoms, const float *d_support, const int64_t *d_action, const float *d_value,
                  const uint8_t *d_terminal, const double *d_is_weight, const double *d_weight_scale, double v_min,
                  double v_max, double discount_n, float *d_target_dist, float *d_loss_row, float *d_abs_error,
                  double *d_loss_elem, float *d_grad_q, void *stream) {
    if (batch < 0 || n_actions < 1 || n_atoms < 2 || n_atoms > kDistMaxAtoms) return MB_ERR_INVALID;
    if (!d_q_dist || !d_next_q_dist || !d_target_next_q_dist || !d_support || !d_action || !d_value || !d_terminal)
        return MB_ERR_INVALID;
    if (batch == 0) return MB_OK;
    // the reference's Python doubles become float32 scalars where they meet a float32 tensor (rainbow.py:247-255)
    DistArgs a{d_q_dist, d_next_q_dist, d_target_next_q_dist, d_support, batch, n_actions, n_atoms, d_action, d_value,
               d_terminal, d_is_weight, d_weight_scale, (float)(v_min + 1e-6), (float)(v_max - 1e-6), (float)v_min,
               (float)((v_max - v_min) / ((double)n_atoms - 1.0)), (float)discount_n, d_target_dist, d_loss_row,
               d_abs_error, d_loss_elem, d_grad_q};
    td_dist_kernel<<<(unsigned)batch, kDistThreads, 0, (cudaStream_t)stream>>>(a);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_td_target(const float *d_next_value, const float *d_reward, const uint8_t *d_terminal, int64_t batch,
                 float discount, const float *d_value, float *d_y, float *d_abs_error, void *stream) {
    if (batch < 0 || !d_next_value || !d_reward || !d_terminal) return MB_ERR_INVALID;
    if (batch == 0) return MB_OK;
    td_target_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_next_value, d_reward, d_terminal, batch, discount, d_value, d_y, d_abs_error);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_soft_update(float *d_target, const float *d_source, int64_t n, double rate, void *stream) {
    if (n < 0 || !d_target || !d_source) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    // python: target * (1.0 - update_rate) + param * update_rate ; both scalars become float32
    soft_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_target, d_source, n, (float)(1.0 - rate), (float)rate);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_soft_update_multi(float *const *d_targets, const float *const *d_sources, const int64_t *sizes,
                         int n_tensors, double rate, void *stream) {
    if (n_tensors < 0 || (n_tensors > 0 && (!d_targets || !d_sources || !sizes))) return MB_ERR_INVALID;
    for (int base = 0; base < n_tensors; base += MB_MAX_SOFT_TENSORS) {
        SoftList l;
        l.n = n_tensors - base < MB_MAX_SOFT_TENSORS ? n_tensors - base : MB_MAX_SOFT_TENSORS;
        l.start[0] = 0;
        for (int k = 0; k < l.n; ++k) {
            if (sizes[base + k] < 0 || !d_targets[base + k] || !d_sources[base + k]) return MB_ERR_INVALID;
            l.target[k] = d_targets[base + k];
            l.source[k] = d_sources[base + k];
            l.start[k + 1] = l.start[k] + sizes[base + k];
        }
        const int64_t total = l.start[l.n];
        if (total == 0) continue;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)mb::kNumSMs * 8) blocks = (int64_t)mb::kNumSMs * 8;
        soft_update_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(l, (float)(1.0 - rate), (float)rate);
        MB_LAUNCHED(1);
    }
    return MB_OK;
}

}  // extern "C"
