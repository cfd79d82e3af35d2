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
#include "mb_common.cuh"

namespace {

constexpr int kEpisodesPerCta = 128;
constexpr int kTile = 2976;  // steps staged per CTA: 4 arrays x (kTile * 33 / 32) floats <= 48 KB

__device__ __forceinline__ int sw(int j) { return j + (j >> 5); }  // de-conflict stride-L access

struct StepTerms {
    float rho, gc, delta;
};

__device__ __forceinline__ StepTerms vtrace_terms(float v, float nv, float r, bool term, float lp, float lmu,
                                                  float g, float clip_c, float clip_rho) {
    const float sim = expf(__fsub_rn(lp, lmu));                    // exp(log_pi - log_mu)       (:322)
    const float c = fminf(clip_c, sim);                            // (:323)
    const float rho = fminf(clip_rho, sim);                        // (:324)
    const float nvm = term ? 0.0f : nv;                            // next_value[terminal] = 0   (:331)
    const float delta = __fmul_rn(rho, __fsub_rn(__fadd_rn(r, __fmul_rn(g, nvm)), v));  // (:332)
    return {rho, __fmul_rn(g, c), delta};
}

__global__ void __launch_bounds__(kEpisodesPerCta) vtrace_kernel(
    const int64_t *__restrict__ offsets, int64_t n_ep, int64_t n_steps, const float *__restrict__ value,
    const float *__restrict__ next_value, const float *__restrict__ reward, const uint8_t *__restrict__ terminal,
    const float *__restrict__ log_pi, const float *__restrict__ log_mu, float g, float clip_c, float clip_rho,
    float *__restrict__ vs_out, float *__restrict__ vss_out, float *__restrict__ rho_out,
    float *__restrict__ adv_out) {
    __shared__ float sV[kTile + kTile / 32 + 1];
    __shared__ float sD[kTile + kTile / 32 + 1];   // delta, then vs
    __shared__ float sG[kTile + kTile / 32 + 1];   // g*c, then vss
    __shared__ float sR[kTile + kTile / 32 + 1];   // reward, then adv
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * kEpisodesPerCta;
    const int64_t e1 = min(e0 + kEpisodesPerCta, n_ep);
    const int64_t fixed_len = offsets ? 0 : n_steps / n_ep;
    const int64_t s0 = offsets ? offsets[e0] : e0 * fixed_len;
    const int64_t s1 = offsets ? offsets[e1] : e1 * fixed_len;
    const int64_t e = e0 + tid;
    int64_t b = 0, len = 0;
    if (e < e1) {
        b = offsets ? offsets[e] : e * fixed_len;
        len = (offsets ? offsets[e + 1] : b + fixed_len) - b;
    }
    if (s1 - s0 <= kTile) {
        const int m = (int)(s1 - s0);
        for (int j = tid; j < m; j += kEpisodesPerCta) {
            const int64_t s = s0 + j;
            const float v = value[s], r = reward[s];
            const StepTerms t = vtrace_terms(v, next_value[s], r, terminal[s] != 0, log_pi[s], log_mu[s], g,
                                             clip_c, clip_rho);
            sV[sw(j)] = v;
            sD[sw(j)] = t.delta;
            sG[sw(j)] = t.gc;
            sR[sw(j)] = r;
            if (rho_out) rho_out[s] = t.rho;
        }
        __syncthreads();
        if (len > 0) {
            const int lb = (int)(b - s0);
            float vs_next = 0.0f, v_next = sV[sw(lb + (int)len - 1)];
            // last step: vs = 0, vss = 0 (:341-357)
            {
                const int j = sw(lb + (int)len - 1);
                const float v = sV[j];
                sD[j] = 0.0f;
                sG[j] = 0.0f;
                sR[j] = __fsub_rn(__fadd_rn(sR[j], __fmul_rn(g, 0.0f)), v);
            }
            for (int k = (int)len - 2; k >= 0; --k) {
                const int j = sw(lb + k);
                const float v = sV[j];
                const float vs = __fadd_rn(__fadd_rn(v, sD[j]), __fmul_rn(sG[j], __fsub_rn(vs_next, v_next)));
                sD[j] = vs;
                sG[j] = vs_next;                                                 // vss[i] = vs[i+1]
                sR[j] = __fsub_rn(__fadd_rn(sR[j], __fmul_rn(g, vs_next)), v);   // r + g*vss - v   (:367)
                vs_next = vs;
                v_next = v;
            }
        }
        __syncthreads();
        for (int j = tid; j < m; j += kEpisodesPerCta) {
            const int64_t s = s0 + j;
            if (vs_out) vs_out[s] = sD[sw(j)];
            if (vss_out) vss_out[s] = sG[sw(j)];
            if (adv_out) adv_out[s] = sR[sw(j)];
        }
    } else if (len > 0) {
        // slice does not fit the staging tile (long episodes): same arithmetic straight from HBM
        float vs_next = 0.0f, v_next = 0.0f;
        for (int64_t k = len - 1; k >= 0; --k) {
            const int64_t s = b + k;
            const float v = value[s], r = reward[s];
            const StepTerms t = vtrace_terms(v, next_value[s], r, terminal[s] != 0, log_pi[s], log_mu[s], g,
                                             clip_c, clip_rho);
            float vs = 0.0f, vss = 0.0f;
            if (k != len - 1) {
                vs = __fadd_rn(__fadd_rn(v, t.delta), __fmul_rn(t.gc, __fsub_rn(vs_next, v_next)));
                vss = vs_next;
            }
            if (rho_out) rho_out[s] = t.rho;
            if (vs_out) vs_out[s] = vs;
            if (vss_out) vss_out[s] = vss;
            if (adv_out) adv_out[s] = __fsub_rn(__fadd_rn(r, __fmul_rn(g, vss)), v);
            vs_next = vs;
            v_next = v;
        }
    }
}

// a2c.py:275-313 in float64 (the reference uses Python floats).  Same three-phase structure as V-trace:
// a CTA owns 64 consecutive episodes (one contiguous slice), streams it into shared memory with coalesced
// loads, runs the dependent recurrence one thread per episode out of shared memory, streams the results out.
constexpr int kGaeEpisodes = 64;
constexpr int kGaeTile = 1664;  // steps staged per CTA: (8 + 8 + 4 + 4 + 1) B per step -> ~42 KB

struct GaeStep {
    double ret, adv;
};

__device__ __forceinline__ GaeStep gae_step(double r, double v, double nv, bool term, bool last, double g, double lam,
                                            double &ret, double &last_v, double &last_a) {
    ret = last ? r : __dadd_rn(__dmul_rn(ret, g), r);                            // value[i-1] = value[i]*g + r
    const double live = 1.0 - (term ? 1.0 : 0.0);
    double a;
    if (lam == 1.0) {
        a = __dsub_rn(ret, v);                                                     // (:284-286)
    } else if (lam == 0.0) {
        a = __dsub_rn(__dadd_rn(r, __dmul_rn(__dmul_rn(g, live), nv)), v);         // (:287-295)
    } else {
        // delta = r + g * last_v * (1 - term) - v ; A = last_a * g * (1 - term) * lam + delta (:296-313)
        const double delta = __dsub_rn(__dadd_rn(r, __dmul_rn(__dmul_rn(g, last_v), live)), v);
        a = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(last_a, g), live), lam), delta);
        last_v = v;
        last_a = a;
    }
    return {ret, a};
}

__global__ void __launch_bounds__(kGaeEpisodes) gae_kernel(const int64_t *__restrict__ offsets, int64_t n_ep,
                                                           int64_t n_steps, const double *__restrict__ reward,
                                                           const uint8_t *__restrict__ terminal,
                                                           const float *__restrict__ cv, const float *__restrict__ ncv,
                                                           double g, double lam, double *__restrict__ ret_out,
                                                           double *__restrict__ gae_out) {
    __shared__ double sR[kGaeTile + kGaeTile / 32 + 1];   // reward, then return
    __shared__ double sA[kGaeTile + kGaeTile / 32 + 1];   // advantage
    __shared__ float sV[kGaeTile + kGaeTile / 32 + 1];
    __shared__ float sN[kGaeTile + kGaeTile / 32 + 1];
    __shared__ uint8_t sT[kGaeTile + kGaeTile / 32 + 1];
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * kGaeEpisodes;
    const int64_t e1 = min(e0 + kGaeEpisodes, n_ep);
    const int64_t fixed_len = offsets ? 0 : n_steps / n_ep;
    const int64_t s0 = offsets ? offsets[e0] : e0 * fixed_len;
    const int64_t s1 = offsets ? offsets[e1] : e1 * fixed_len;
    const int64_t e = e0 + tid;
    int64_t b = 0, len = 0;
    if (e < e1) {
        b = offsets ? offsets[e] : e * fixed_len;
        len = (offsets ? offsets[e + 1] : b + fixed_len) - b;
    }
    const bool need_next = gae_out && lam == 0.0;
    if (s1 - s0 <= kGaeTile) {
        const int m = (int)(s1 - s0);
        for (int j = tid; j < m; j += kGaeEpisodes) {
            const int64_t s = s0 + j;
            sR[sw(j)] = reward[s];
            if (gae_out) {
                sV[sw(j)] = cv[s];
                sT[sw(j)] = terminal[s];
                if (need_next) sN[sw(j)] = ncv[s];
            }
        }
        __syncthreads();
        if (len > 0) {
            const int lb = (int)(b - s0);
            double ret = 0.0, last_v = 0.0, last_a = 0.0;
            for (int k = (int)len - 1; k >= 0; --k) {
                const int j = sw(lb + k);
                const GaeStep o = gae_step(sR[j], gae_out ? (double)sV[j] : 0.0, need_next ? (double)sN[j] : 0.0,
                                           gae_out ? sT[j] != 0 : false, k == (int)len - 1, g, lam, ret, last_v, last_a);
                sR[j] = o.ret;
                sA[j] = o.adv;
            }
        }
        __syncthreads();
        for (int j = tid; j < m; j += kGaeEpisodes) {
            const int64_t s = s0 + j;
            if (ret_out) ret_out[s] = sR[sw(j)];
            if (gae_out) gae_out[s] = sA[sw(j)];
        }
    } else if (len > 0) {
        double ret = 0.0, last_v = 0.0, last_a = 0.0;
        for (int64_t k = len - 1; k >= 0; --k) {
            const int64_t s = b + k;
            const GaeStep o = gae_step(reward[s], gae_out ? (double)cv[s] : 0.0, need_next ? (double)ncv[s] : 0.0,
                                       gae_out ? terminal[s] != 0 : false, k == len - 1, g, lam, ret, last_v, last_a);
            if (ret_out) ret_out[s] = o.ret;
            if (gae_out) gae_out[s] = o.adv;
        }
    }
}

}  // namespace

extern "C" {

int mb_vtrace(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const float *d_value,
              const float *d_next_value, const float *d_reward, const uint8_t *d_terminal, const float *d_log_pi,
              const float *d_log_mu, float discount, float clip_c, float clip_rho, float *d_vs, float *d_vss,
              float *d_rho, float *d_adv, void *stream) {
    if (n_episodes < 0 || n_steps < 0) return MB_ERR_INVALID;
    if (n_episodes == 0 || n_steps == 0) return MB_OK;
    if (!d_value || !d_next_value || !d_reward || !d_terminal || !d_log_pi || !d_log_mu) return MB_ERR_INVALID;
    if (!d_offsets && n_steps % n_episodes != 0) return MB_ERR_INVALID;
    const unsigned blocks = (unsigned)((n_episodes + kEpisodesPerCta - 1) / kEpisodesPerCta);
    vtrace_kernel<<<blocks, kEpisodesPerCta, 0, (cudaStream_t)stream>>>(
        d_offsets, n_episodes, n_steps, d_value, d_next_value, d_reward, d_terminal, d_log_pi, d_log_mu, discount,
        clip_c, clip_rho, d_vs, d_vss, d_rho, d_adv);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_gae(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const double *d_reward,
           const uint8_t *d_terminal, const float *d_critic_value, const float *d_next_critic_value,
           double discount, double gae_lambda, double *d_returns, double *d_gae, void *stream) {
    if (n_episodes < 0 || n_steps < 0) return MB_ERR_INVALID;
    if (n_episodes == 0 || n_steps == 0) return MB_OK;
    if (!d_reward) return MB_ERR_INVALID;
    if (d_gae && (!d_terminal || !d_critic_value)) return MB_ERR_INVALID;
    if (d_gae && gae_lambda == 0.0 && !d_next_critic_value) return MB_ERR_INVALID;
    if (!d_offsets && n_steps % n_episodes != 0) return MB_ERR_INVALID;
    gae_kernel<<<(unsigned)((n_episodes + kGaeEpisodes - 1) / kGaeEpisodes), kGaeEpisodes, 0, (cudaStream_t)stream>>>(
        d_offsets, n_episodes, n_steps, d_reward, d_terminal, d_critic_value, d_next_critic_value, discount,
        gae_lambda, d_returns, d_gae);
    MB_LAUNCHED(1);
    return MB_OK;
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
    r = mb_vtrace(h_offsets ? (const int64_t *)d : nullptr, n_episodes, n_steps, (const float *)(d + o_v),
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
    r = mb_gae(h_offsets ? (const int64_t *)d : nullptr, n_episodes, n_steps, (const double *)(d + o_r),
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
