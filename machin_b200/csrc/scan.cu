// Trajectory reverse scans for sm_100a: IMPALA V-trace and A2C/A3C returns + GAE.
//
// Replaces the per-step Python loops of impala.py:322-362 (scalar torch ops, forced onto the CPU at
// :318-330) and a2c.py:275-313 (Python floats with one critic forward per step).
//
// V-trace: a CTA owns 128 consecutive episodes, i.e. one contiguous slice of the flattened
// episode-major step arrays.  Phase A streams the slice in with coalesced loads and evaluates the
// independent per-step terms (rho, c, delta) into shared memory; phase B runs the dependent
// recurrence, one thread per episode, entirely out of shared memory; phase C streams vs / vss / adv
// back out coalesced.  Each float32 operation is rounded separately in the reference's order
// ((v + delta) + (g*c) * (vs' - v')), so apart from exp() the results are bit-identical to CPU torch.
#include <algorithm>

#include "mb_common.cuh"

namespace {

#include "scan_device.cuh"

}  // namespace

extern "C" {

int mb_vtrace_len(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, int64_t max_episode_len,
                  const float *d_value, const float *d_next_value, const float *d_reward, const uint8_t *d_terminal,
                  const float *d_log_pi, const float *d_log_mu, float discount, float clip_c, float clip_rho, float *d_vs,
                  float *d_vss, float *d_rho, float *d_adv, void *stream) {
    if (n_episodes < 0 || n_steps < 0) return MB_ERR_INVALID;
    if (n_episodes == 0 || n_steps == 0) return MB_OK;
    if (!d_value || !d_next_value || !d_reward || !d_terminal || !d_log_pi || !d_log_mu) return MB_ERR_INVALID;
    if (!d_offsets && n_steps % n_episodes != 0) return MB_ERR_INVALID;
    if (!d_offsets) max_episode_len = n_steps / n_episodes;
    const int epc = episodes_per_cta(max_episode_len, kTile, kEpisodesPerCta);
    const unsigned blocks = (unsigned)((n_episodes + epc - 1) / epc);
    vtrace_kernel<<<blocks, kEpisodesPerCta, 0, (cudaStream_t)stream>>>(
        d_offsets, n_episodes, n_steps, d_value, d_next_value, d_reward, d_terminal, d_log_pi, d_log_mu, discount,
        clip_c, clip_rho, d_vs, d_vss, d_rho, d_adv, epc);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_vtrace(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const float *d_value,
              const float *d_next_value, const float *d_reward, const uint8_t *d_terminal, const float *d_log_pi,
              const float *d_log_mu, float discount, float clip_c, float clip_rho, float *d_vs, float *d_vss,
              float *d_rho, float *d_adv, void *stream) {
    return mb_vtrace_len(d_offsets, n_episodes, n_steps, 0, d_value, d_next_value, d_reward, d_terminal, d_log_pi, d_log_mu,
                         discount, clip_c, clip_rho, d_vs, d_vss, d_rho, d_adv, stream);
}

int mb_gae_len(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, int64_t max_episode_len,
               const double *d_reward, const uint8_t *d_terminal, const float *d_critic_value,
               const float *d_next_critic_value, double discount, double gae_lambda, double *d_returns, double *d_gae,
               void *stream) {
    if (n_episodes < 0 || n_steps < 0) return MB_ERR_INVALID;
    if (n_episodes == 0 || n_steps == 0) return MB_OK;
    if (!d_reward) return MB_ERR_INVALID;
    if (d_gae && (!d_terminal || !d_critic_value)) return MB_ERR_INVALID;
    if (d_gae && gae_lambda == 0.0 && !d_next_critic_value) return MB_ERR_INVALID;
    if (!d_offsets && n_steps % n_episodes != 0) return MB_ERR_INVALID;
    if (!d_offsets) max_episode_len = n_steps / n_episodes;
    const int epc = episodes_per_cta(max_episode_len, kGaeTile, kGaeEpisodes);
    gae_kernel<<<(unsigned)((n_episodes + epc - 1) / epc), kGaeEpisodes, 0, (cudaStream_t)stream>>>(
        d_offsets, n_episodes, n_steps, d_reward, d_terminal, d_critic_value, d_next_critic_value, discount,
        gae_lambda, d_returns, d_gae, epc);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_gae(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const double *d_reward,
           const uint8_t *d_terminal, const float *d_critic_value, const float *d_next_critic_value,
           double discount, double gae_lambda, double *d_returns, double *d_gae, void *stream) {
    return mb_gae_len(d_offsets, n_episodes, n_steps, 0, d_reward, d_terminal, d_critic_value, d_next_critic_value, discount,
                      gae_lambda, d_returns, d_gae, stream);
}

// ------------------------------------------------------------------------------ host variants
int mb_vtrace_host(int device, const int64_t *h_offsets, int64_t n_episodes, int64_t n_steps, const float *h_value,
                   const float *h_next_value, const float *h_reward, const uint8_t *h_terminal,
                   const float *h_log_pi, const float *h_log_mu, float discount, float clip_c, float clip_rho,
                   float *h_vs, float *h_vss, float *h_rho, float *h_adv) {
    if (n_episodes <= 0 || n_steps <= 0) return n_episodes < 0 || n_steps < 0 ? MB_ERR_INVALID : MB_OK;
    if (mb_device_count() <= device || device < 0) return MB_ERR_NO_DEVICE;
    mb::DeviceGuard g(device);
    mb::Staging *st = mb::staging_for(device);
    if (!st) return MB_ERR_CUDA;
    std::lock_guard<std::mutex> staging_lock(st->mu);
    const size_t S = (size_t)n_steps, fb = 4 * S;
    const size_t off_bytes = h_offsets ? 8 * (size_t)(n_episodes + 1) : 0;
    const size_t in_bytes = off_bytes + 5 * fb + ((S + 15) / 16) * 16;
    const size_t total = in_bytes + 4 * fb;
    int r = st->ensure(total, total);
    if (r != MB_OK) return r;
    char *h = st->h;
    size_t o = 0;
    if (h_offsets) { memcpy(h, h_offsets, off_bytes); o = off_bytes; }
    const size_t o_v = o, o_nv = o + fb, o_r = o + 2 * fb, o_lp = o + 3 * fb, o_lm = o + 4 * fb, o_t = o + 5 * fb;
    memcpy(h + o_v, h_value, fb);
    memcpy(h + o_nv, h_next_value, fb);
    memcpy(h + o_r, h_reward, fb);
    memcpy(h + o_lp, h_log_pi, fb);
    memcpy(h + o_lm, h_log_mu, fb);
    memcpy(h + o_t, h_terminal, S);
    MB_CUDA(cudaMemcpyAsync(st->d, h, in_bytes, cudaMemcpyHostToDevice, st->stream));
    char *d = st->d;
    float *d_vs = (float *)(d + in_bytes), *d_vss = d_vs + S, *d_rho = d_vss + S, *d_adv = d_rho + S;
    int64_t longest = 0;   // the host knows the episode lengths: keep the staged path for long episodes too
    if (h_offsets)
        for (int64_t e = 0; e < n_episodes; ++e) longest = std::max(longest, h_offsets[e + 1] - h_offsets[e]);
    r = mb_vtrace_len(h_offsets ? (const int64_t *)d : nullptr, n_episodes, n_steps, longest, (const float *)(d + o_v),
                  (const float *)(d + o_nv), (const float *)(d + o_r), (const uint8_t *)(d + o_t),
                  (const float *)(d + o_lp), (const float *)(d + o_lm), discount, clip_c, clip_rho, d_vs, d_vss,
                  d_rho, d_adv, st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaMemcpyAsync(h + in_bytes, d + in_bytes, 4 * fb, cudaMemcpyDeviceToHost, st->stream));
    MB_CUDA(cudaStreamSynchronize(st->stream));
    if (h_vs) memcpy(h_vs, h + in_bytes, fb);
    if (h_vss) memcpy(h_vss, h + in_bytes + fb, fb);
    if (h_rho) memcpy(h_rho, h + in_bytes + 2 * fb, fb);
    if (h_adv) memcpy(h_adv, h + in_bytes + 3 * fb, fb);
    return MB_OK;
}

int mb_gae_host(int device, const int64_t *h_offsets, int64_t n_episodes, int64_t n_steps, const double *h_reward,
                const uint8_t *h_terminal, const float *h_critic_value, const float *h_next_critic_value,
                double discount, double gae_lambda, double *h_returns, double *h_gae) {
    if (n_episodes <= 0 || n_steps <= 0) return n_episodes < 0 || n_steps < 0 ? MB_ERR_INVALID : MB_OK;
    if (mb_device_count() <= device || device < 0) return MB_ERR_NO_DEVICE;
    if (!h_reward || !h_terminal || !h_critic_value) return MB_ERR_INVALID;
    mb::DeviceGuard g(device);
    mb::Staging *st = mb::staging_for(device);
    if (!st) return MB_ERR_CUDA;
    std::lock_guard<std::mutex> staging_lock(st->mu);
    const size_t S = (size_t)n_steps;
    const size_t off_bytes = h_offsets ? 8 * (size_t)(n_episodes + 1) : 0;
    const size_t o_r = off_bytes, o_cv = o_r + 8 * S, o_ncv = o_cv + 4 * S, o_t = o_ncv + 4 * S;
    const size_t in_bytes = ((o_t + S + 15) / 16) * 16;
    const size_t total = in_bytes + 16 * S;
    int r = st->ensure(total, total);
    if (r != MB_OK) return r;
    char *h = st->h;
    if (h_offsets) memcpy(h, h_offsets, off_bytes);
    memcpy(h + o_r, h_reward, 8 * S);
    memcpy(h + o_cv, h_critic_value, 4 * S);
    if (h_next_critic_value) memcpy(h + o_ncv, h_next_critic_value, 4 * S);
    else memset(h + o_ncv, 0, 4 * S);
    memcpy(h + o_t, h_terminal, S);
    MB_CUDA(cudaMemcpyAsync(st->d, h, in_bytes, cudaMemcpyHostToDevice, st->stream));
    char *d = st->d;
    double *d_ret = (double *)(d + in_bytes), *d_adv = d_ret + S;
    int64_t longest = 0;
    if (h_offsets)
        for (int64_t e = 0; e < n_episodes; ++e) longest = std::max(longest, h_offsets[e + 1] - h_offsets[e]);
    r = mb_gae_len(h_offsets ? (const int64_t *)d : nullptr, n_episodes, n_steps, longest, (const double *)(d + o_r),
               (const uint8_t *)(d + o_t), (const float *)(d + o_cv), (const float *)(d + o_ncv), discount,
               gae_lambda, d_ret, d_adv, st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaMemcpyAsync(h + in_bytes, d + in_bytes, 16 * S, cudaMemcpyDeviceToHost, st->stream));
    MB_CUDA(cudaStreamSynchronize(st->stream));
    if (h_returns) memcpy(h_returns, h + in_bytes, 8 * S);
    if (h_gae) memcpy(h_gae, h + in_bytes + 8 * S, 8 * S);
    return MB_OK;
}

}  // extern "C"
