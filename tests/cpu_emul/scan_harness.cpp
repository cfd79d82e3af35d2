// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/scan_device.cuh (V-trace, returns + GAE reverse scans) for
// the host through cuda_shim.h; C entry points for tests/test_scan_kernels_emulated_cpu.py with the launch
// geometry of csrc/scan.cu (one CTA per 128 / 64 consecutive episodes).
#include "cuda_shim.h"

#include <algorithm>
using std::min;
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }

#include "scan_device.cuh"

extern "C" {

int emul_vtrace(const int64_t *offsets, int64_t n_ep, int64_t n_steps, const float *value, const float *next_value,
                const float *reward, const uint8_t *terminal, const float *log_pi, const float *log_mu, float g,
                float clip_c, float clip_rho, float *vs, float *vss, float *rho, float *adv, int64_t max_len) {
    if (!offsets) max_len = n_steps / n_ep;
    const int epc = episodes_per_cta(max_len, kTile, kEpisodesPerCta);      // same choice as csrc/scan.cu
    shim_launch((unsigned)((n_ep + epc - 1) / epc), kEpisodesPerCta, [&] {
        vtrace_kernel(offsets, n_ep, n_steps, value, next_value, reward, terminal, log_pi, log_mu, g, clip_c, clip_rho, vs,
                      vss, rho, adv, epc);
    });
    return 0;
}

int emul_gae(const int64_t *offsets, int64_t n_ep, int64_t n_steps, const double *reward, const uint8_t *terminal,
             const float *cv, const float *ncv, double g, double lam, double *ret, double *gae, int64_t max_len) {
    if (!offsets) max_len = n_steps / n_ep;
    const int epc = episodes_per_cta(max_len, kGaeTile, kGaeEpisodes);
    shim_launch((unsigned)((n_ep + epc - 1) / epc), kGaeEpisodes, [&] {
        gae_kernel(offsets, n_ep, n_steps, reward, terminal, cv, ncv, g, lam, ret, gae, epc);
    });
    return 0;
}

}  // extern "C"
