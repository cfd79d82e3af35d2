// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/sumtree_sample.cuh (sum-tree descent, fused PER draw)
// for the host through cuda_shim_mp.h.  Usage: sample_emul <in.bin> <out.bin>
//   in : int64 header [variant, size, depth, n, R, n_stored, n_draw], f64 beta[R], f64 all_weight_sum,
//        f64 W[2^depth - 1], f64 u[R * n]   (variant 0: u holds n query weights)
//   out: i64 index[R * n], f64 is_weight[R * n]
// variant: 0 tree_find_kernel, 1 per_sample_small, 2 per_sample_cluster, 3 per_sample_cluster_multi (R requesters,
//          device-side parameter blocks), 4 per_sample_cluster with a device-side parameter block (n_draw < n)
#include "cuda_shim_mp.h"

#include "machin_b200.h"
#include "sumtree_update.cuh"   // level_off
#include "sumtree_sample.cuh"

template <typename T> static T *shared_copy(const T *src, size_t n) {
    T *p = static_cast<T *>(shim::shared_alloc(sizeof(T) * (n ? n : 1)));
    if (n && src) std::memcpy(p, src, sizeof(T) * n);
    return p;
}

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t hd[7];
    if (std::fread(hd, 8, 7, f) != 7) return 2;
    const int variant = (int)hd[0], depth = (int)hd[2], R = (int)hd[4];
    const int64_t size = hd[1], n = hd[3], n_stored = hd[5], n_draw = hd[6], nodes = ((int64_t)1 << depth) - 1;
    std::vector<double> beta(R), W(nodes), u((size_t)R * n);
    double all_sum;
    if (std::fread(beta.data(), 8, R, f) != (size_t)R || std::fread(&all_sum, 8, 1, f) != 1 ||
        std::fread(W.data(), 8, nodes, f) != (size_t)nodes || std::fread(u.data(), 8, u.size(), f) != u.size())
        return 2;
    std::fclose(f);
    TreeDev t{};
    t.W = shared_copy(W.data(), nodes);
    TreeState st{};
    t.st = shared_copy(&st, 1);
    t.size = size;
    t.total = (int64_t)1 << depth;
    t.depth = depth;
    const double *d_u = shared_copy(u.data(), u.size());
    int64_t *d_idx = shared_copy<int64_t>(nullptr, (size_t)R * n);
    double *d_isw = shared_copy<double>(nullptr, (size_t)R * n);
    std::vector<mb_per_params> prm(R);
    for (int r = 0; r < R; ++r) prm[r] = mb_per_params{beta[r], all_sum, n_stored, 0, n_draw};
    const mb_per_params *d_prm = shared_copy(prm.data(), R);
    const unsigned threads = n <= 256 ? 32 : (n <= 512 ? 64 : 128);
    const int ctas = (int)((n + threads - 1) / threads);
    int rc = 9;
    if (variant == 0) {
        rc = shim::launch((int)((n + 255) / 256), 256, 0, [=] { tree_find_kernel(t, d_u, n, d_idx); });
    } else if (variant == 1) {
        PerArgs a{n_stored, all_sum, beta[0], nullptr};
        rc = shim::launch(1, (unsigned)((n + 31) / 32 * 32), 0, [=] { per_sample_small(t, d_u, n, a, d_idx, d_isw); });
    } else if (variant == 2) {
        PerArgs a{n_stored, all_sum, beta[0], nullptr};
        rc = shim::launch(ctas, threads, 0, [=] { per_sample_cluster(t, d_u, n, a, d_idx, d_isw); });
    } else if (variant == 3) {
        rc = shim::launch(ctas, threads, 0, [=] { per_sample_cluster_multi(t, d_u, n, d_prm, d_idx, d_isw); }, R);
    } else if (variant == 4) {
        PerArgs a{0, 0.0, 0.0, d_prm};
        rc = shim::launch(ctas, threads, 0, [=] { per_sample_cluster(t, d_u, n, a, d_idx, d_isw); });
    }
    if (rc) return rc;
    f = std::fopen(argv[2], "wb");
    if (!f) return 2;
    std::fwrite(d_idx, 8, (size_t)R * n, f);
    std::fwrite(d_isw, 8, (size_t)R * n, f);
    std::fclose(f);
    return 0;
}
