// IMPALA V-trace and A2C/A3C returns + GAE reverse scans: device code only (no CUDA runtime calls).  Included by
// scan.cu inside its anonymous namespace and by tests/cpu_emul/scan_harness.cpp, which compiles THIS SOURCE with
// g++ against the CUDA shim.  See scan.cu for the design.
#pragma once
#include <stdint.h>

constexpr int kEpisodesPerCta = 128;
constexpr int kTile = 2976;  // steps staged per CTA: 4 arrays x (kTile * 33 / 32) floats <= 48 KB

// episodes per CTA for a given longest episode (0 = unknown): as many as fit the staging tile, at most `cap`
inline int episodes_per_cta(int64_t max_len, int tile, int cap) {
    if (max_len <= 0) return cap;
    const int64_t e = tile / max_len;
    return (int)(e < 1 ? 1 : (e > cap ? cap : e));
}

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
    float *__restrict__ adv_out, int epc) {
    __shared__ float sV[kTile + kTile / 32 + 1];
    __shared__ float sD[kTile + kTile / 32 + 1];   // delta, then vs
    __shared__ float sG[kTile + kTile / 32 + 1];   // g*c, then vss
    __shared__ float sR[kTile + kTile / 32 + 1];   // reward, then adv
    const int tid = threadIdx.x;
    // epc (<= blockDim.x) episodes per CTA: the host shrinks it for long episodes so that a CTA's slice still fits
    // the staging tile (128 episodes of more than 23 steps would not) and the coalesced three-phase path is kept
    const int64_t e0 = (int64_t)blockIdx.x * epc;
    const int64_t e1 = min(e0 + epc, n_ep);
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
    } else if (epc == 1) {
        // ONE episode longer than the staging tile: walk it tile by tile from its end -- coalesced load of the tile,
        // thread 0 runs the recurrence out of shared memory carrying (vs_next, v_next) across tiles, coalesced store
        float vs_next = 0.0f, v_next = 0.0f;
        for (int64_t hi = s1; hi > s0; hi -= kTile) {
            const int64_t lo = hi - kTile > s0 ? hi - kTile : s0;
            const int m = (int)(hi - lo);
            for (int j = tid; j < m; j += kEpisodesPerCta) {
                const int64_t s = lo + j;
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
            if (tid == 0) {
                for (int k = m - 1; k >= 0; --k) {
                    const int j = sw(k);
                    const float v = sV[j];
                    float vs = 0.0f, vss = 0.0f;
                    if (lo + k != s1 - 1) {                                          // last step: vs = vss = 0 (:341-357)
                        vs = __fadd_rn(__fadd_rn(v, sD[j]), __fmul_rn(sG[j], __fsub_rn(vs_next, v_next)));
                        vss = vs_next;
                    }
                    sD[j] = vs;
                    sG[j] = vss;
                    sR[j] = __fsub_rn(__fadd_rn(sR[j], __fmul_rn(g, vss)), v);       // r + g*vss - v   (:367)
                    vs_next = vs;
                    v_next = v;
                }
            }
            __syncthreads();
            for (int j = tid; j < m; j += kEpisodesPerCta) {
                const int64_t s = lo + j;
                if (vs_out) vs_out[s] = sD[sw(j)];
                if (vss_out) vss_out[s] = sG[sw(j)];
                if (adv_out) adv_out[s] = sR[sw(j)];
            }
            __syncthreads();                                                         // the next tile overwrites the staging arrays
        }
    } else if (len > 0) {
        // slice does not fit the staging tile and the launch did not know the lengths: same arithmetic, one thread
        // per episode straight from HBM
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
                                                           double *__restrict__ gae_out, int epc) {
    __shared__ double sR[kGaeTile + kGaeTile / 32 + 1];   // reward, then return
    __shared__ double sA[kGaeTile + kGaeTile / 32 + 1];   // advantage
    __shared__ float sV[kGaeTile + kGaeTile / 32 + 1];
    __shared__ float sN[kGaeTile + kGaeTile / 32 + 1];
    __shared__ uint8_t sT[kGaeTile + kGaeTile / 32 + 1];
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * epc;         // epc <= blockDim.x episodes per CTA, see vtrace_kernel
    const int64_t e1 = min(e0 + epc, n_ep);
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
    } else if (epc == 1) {
        // ONE episode longer than the staging tile: tile by tile from its end, (ret, last_v, last_a) carried by thread 0
        double ret = 0.0, last_v = 0.0, last_a = 0.0;
        for (int64_t hi = s1; hi > s0; hi -= kGaeTile) {
            const int64_t lo = hi - kGaeTile > s0 ? hi - kGaeTile : s0;
            const int m = (int)(hi - lo);
            for (int j = tid; j < m; j += kGaeEpisodes) {
                const int64_t s = lo + j;
                sR[sw(j)] = reward[s];
                if (gae_out) {
                    sV[sw(j)] = cv[s];
                    sT[sw(j)] = terminal[s];
                    if (need_next) sN[sw(j)] = ncv[s];
                }
            }
            __syncthreads();
            if (tid == 0) {
                for (int k = m - 1; k >= 0; --k) {
                    const int j = sw(k);
                    const GaeStep o = gae_step(sR[j], gae_out ? (double)sV[j] : 0.0, need_next ? (double)sN[j] : 0.0,
                                               gae_out ? sT[j] != 0 : false, lo + k == s1 - 1, g, lam, ret, last_v, last_a);
                    sR[j] = o.ret;
                    sA[j] = o.adv;
                }
            }
            __syncthreads();
            for (int j = tid; j < m; j += kGaeEpisodes) {
                const int64_t s = lo + j;
                if (ret_out) ret_out[s] = sR[sw(j)];
                if (gae_out) gae_out[s] = sA[sw(j)];
            }
            __syncthreads();
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
