// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/sumtree_update.cuh (the device source of the small-batch
// sum-tree update kernels) for the host through cuda_shim_mp.h.  Usage: tree_emul <in.bin> <out.bin>
//   in : int64 header [variant, size, depth, n, dtype(0 f32 / 1 f64), normalize, has_valid] , f64 eps, f64 alpha,
//        f64 max_leaf, f64 W[2^depth - 1], i64 idx[n], (f32|f64) w[n], u8 valid[n] (if has_valid)
//   out: f64 W[2^depth - 1], f64 max_leaf, i64 max_is_f32, i64 err
// variant: 0 tree_update_sorted (GPU-validated; checks the shim), 1 v2<CB=0, async>, 2 v2<CB=3, barrier>,
//          3 v2<CB=3, async>, 4 tree_update_kernel<.., false> (generic single-CTA body),
//          5 tree_update_part (one CTA per subtree, CTAs run one after the other: the kernel has no cross-CTA wait)
#include "cuda_shim_mp.h"

#include "machin_b200.h"
#include "sumtree_update.cuh"

template <typename T> static T *shared_copy(const T *src, size_t n) {
    T *p = static_cast<T *>(shim::shared_alloc(sizeof(T) * (n ? n : 1)));
    if (n) std::memcpy(p, src, sizeof(T) * n);
    return p;
}

template <typename WT>
static int run(int variant, TreeDev t, const int64_t *idx, const WT *w, const uint8_t *valid, int n, Normalize nz) {
    const unsigned threads = (unsigned)((n + 31) / 32 * 32);
    const size_t dyn = sizeof(double) * (size_t)kSortedN * (size_t)(t.depth - 1);
    switch (variant) {
    case 0: return shim::launch(1, threads, dyn, [=] { tree_update_sorted<WT>(t, idx, w, valid, n, nz, 1); });
    case 1: return shim::launch(1, threads, dyn, [=] { tree_update_sorted_v2<WT, 0, true>(t, idx, w, valid, n, nz, 1); });
    case 2: return shim::launch(8, threads, dyn, [=] { tree_update_sorted_v2<WT, 3, false>(t, idx, w, valid, n, nz, 1); });
    case 3: return shim::launch(8, threads, dyn, [=] { tree_update_sorted_v2<WT, 3, true>(t, idx, w, valid, n, nz, 1); });
    case 4: return shim::launch(1, threads, 0, [=] { tree_update_kernel<WT, false>(t, idx, w, valid, (int64_t)n, nz, 1); });
    case 5: {
        const int levels = t.depth - 1;
        return shim::launch(1, kPartThreads, part_dyn_smem(levels),
                            [=] { tree_update_part<WT>(t, idx, w, valid, n, nz, 1); }, 1 << part_cut(levels));
    }
    }
    return 9;
}

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t hd[7];
    double eps, alpha, max_leaf;
    if (std::fread(hd, 8, 7, f) != 7 || std::fread(&eps, 8, 1, f) != 1 || std::fread(&alpha, 8, 1, f) != 1 ||
        std::fread(&max_leaf, 8, 1, f) != 1)
        return 2;
    const int variant = (int)hd[0], depth = (int)hd[2], n = (int)hd[3];
    const int64_t size = hd[1], total = (int64_t)1 << depth, nodes = total - 1;
    std::vector<double> W(nodes);
    std::vector<int64_t> idx(n);
    std::vector<unsigned char> wraw((size_t)n * (hd[4] ? 8 : 4)), valid(hd[6] ? n : 0);
    if (std::fread(W.data(), 8, nodes, f) != (size_t)nodes || std::fread(idx.data(), 8, n, f) != (size_t)n ||
        std::fread(wraw.data(), 1, wraw.size(), f) != wraw.size() ||
        (hd[6] && std::fread(valid.data(), 1, n, f) != (size_t)n))
        return 2;
    std::fclose(f);
    TreeDev t{};
    t.W = shared_copy(W.data(), nodes);
    std::vector<int> owner(size, -1);
    t.owner = shared_copy(owner.data(), size);
    TreeState st{};
    st.max_leaf = max_leaf;
    st.batch_max = -INFINITY;
    t.st = shared_copy(&st, 1);
    t.size = size;
    t.total = total;
    t.depth = depth;
    const int64_t *d_idx = shared_copy(idx.data(), n);
    const uint8_t *d_valid = hd[6] ? shared_copy(valid.data(), n) : nullptr;
    Normalize nz{(int)hd[5], eps, alpha};
    int rc;
    if (hd[4]) rc = run<double>(variant, t, d_idx, shared_copy(reinterpret_cast<double *>(wraw.data()), n), d_valid, n, nz);
    else rc = run<float>(variant, t, d_idx, shared_copy(reinterpret_cast<float *>(wraw.data()), n), d_valid, n, nz);
    if (rc) return rc;
    f = std::fopen(argv[2], "wb");
    if (!f) return 2;
    const int64_t is_f32 = t.st->max_is_f32, err = t.st->err;
    std::fwrite(t.W, 8, nodes, f);
    std::fwrite(&t.st->max_leaf, 8, 1, f);
    std::fwrite(&is_f32, 8, 1, f);
    std::fwrite(&err, 8, 1, f);
    std::fclose(f);
    return 0;
}
