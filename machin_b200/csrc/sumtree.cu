// Device-resident sum tree + fused PER draw for sm_100a.
//
// Replaces machin/frame/buffers/prioritized_buffer.py:8-231 (WeightTree) and :397-434
// (PrioritizedBuffer.sample_index_and_weight / _normalize_priority) on the learner hot path.
//
// HBM layout: ONE float64 array in the reference's flattened order -- level 0 (leaves, padded
// to 2^(depth-1)) first, root last.  With T = 2^depth, level l starts at T - (T >> l) and has
// T >> (l+1) nodes, so every level start is a multiple of 2^(depth-l) elements: any 2^J-child
// group is naturally 8*2^J-byte aligned and a 16-child group is exactly one 128-byte line.
//
// Invariant kept by every mutator here (and by every mutator of the reference): each parent
// holds fl(left + right) of its stored children.  The descent kernel relies on it to jump
// several levels per dependent load: it fetches the 2^J descendants of the current node and
// re-derives the intermediate sums with the same pairwise additions, which reproduces the
// stored intermediate nodes bit-for-bit, then applies the reference's per-level comparisons.
#include <cooperative_groups.h>

#include <stdlib.h>

#include <new>

#include "mb_common.cuh"

namespace cg = cooperative_groups;

#include "sumtree_types.cuh"   // TreeState, TreeDev

struct mb_tree {
    int device;
    TreeDev dev;
    int64_t n_leaf_pad;
    int64_t n_nodes;
    TreeState *h_state;  // pinned mirror for mb_tree_stats / mb_tree_check
    int coop_blocks;     // co-resident CTAs for the cooperative update kernel
    int gen;             // update-launch generation counter
    int sorted_cluster;  // MB_SORTED_CLUSTER=1: small sorted updates on a cluster of 8 CTAs (opt-in until measured)
    int sorted_async;    // MB_SORTED_ASYNC=1: barrier-free walk (point-to-point hand-overs), opt-in until measured
    int legacy_update;   // MB_UPDATE_LEGACY=1: the single-CTA / cluster sorted kernels instead of tree_update_part (A/B timing)
    int force_generic;   // MB_FORCE_GENERIC_UPDATE=1: single-CTA PER draw instead of the cluster kernel (A/B timing)
};

namespace {

#include "sumtree_update.cuh"   // level_off, atomic_max_f64, the small-batch update kernels

#include "sumtree_sample.cuh"   // descent (find), fused PER draw

// ---------------------------------------------------------------------------------------------
// build  (update_all_leaves + _build, prioritized_buffer.py:188-199, 216-231)
// ---------------------------------------------------------------------------------------------
__global__ void tree_state_reset(TreeState *st, double max_leaf) {
    st->max_leaf = max_leaf;
    st->batch_max = -INFINITY;
    st->max_is_f32 = 0;
}

// One pass lifts 11 levels.  CTA = 256 threads x 8 source nodes = 2048 nodes of level `lv`.
// Reads 8 B/node once, writes every produced level once: ~16 B per source node in total.
__global__ void __launch_bounds__(256) tree_build_pass(TreeDev t, int lv, const double *src_leaves) {
    __shared__ double warp_tot[8];
    __shared__ double warp_mx[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t gthread = (int64_t)blockIdx.x * 256 + tid;
    const int64_t base = gthread * 8;
    double *src = t.W + level_off(t, lv) + base;
    double v[8];
    if (lv == 0 && src_leaves != nullptr) {
        double mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t i = base + k;
            v[k] = (i < t.size) ? __ldg(src_leaves + i) : src[k];
            if (i < t.size) mx = fmax(mx, v[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2 *>(src + k) = make_double2(v[k], v[k + 1]);
        mx = mb::warp_max(mx);
        if (lane == 0) warp_mx[wid] = mx;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            double2 a = __ldcg(reinterpret_cast<const double2 *>(src + k));
            v[k] = a.x;
            v[k + 1] = a.y;
        }
        if (lv == 0) {
            double mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (base + k < t.size) mx = fmax(mx, v[k]);
            mx = mb::warp_max(mx);
            if (lane == 0) warp_mx[wid] = mx;
        }
    }
    // levels lv+1 .. lv+3 inside the thread
    double a1[4], a2[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) a1[k] = v[2 * k] + v[2 * k + 1];
    a2[0] = a1[0] + a1[1];
    a2[1] = a1[2] + a1[3];
    double x = a2[0] + a2[1];
    double *l1 = t.W + level_off(t, lv + 1) + gthread * 4;
    *reinterpret_cast<double2 *>(l1) = make_double2(a1[0], a1[1]);
    *reinterpret_cast<double2 *>(l1 + 2) = make_double2(a1[2], a1[3]);
    *reinterpret_cast<double2 *>(t.W + level_off(t, lv + 2) + gthread * 2) = make_double2(a2[0], a2[1]);
    t.W[level_off(t, lv + 3) + gthread] = x;
    // levels lv+4 .. lv+8 across the warp (a + b == b + a bitwise, so every lane gets the node)
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        x = x + __shfl_xor_sync(0xffffffffu, x, 1 << s);
        if ((lane & ((2 << s) - 1)) == 0) t.W[level_off(t, lv + 4 + s) + (gthread >> (s + 1))] = x;
    }
    if (lane == 0) warp_tot[wid] = x;
    __syncthreads();
    if (tid == 0) {
        // levels lv+9 .. lv+11 across the 8 warps
        const int64_t cta = blockIdx.x;
        double b0 = warp_tot[0] + warp_tot[1], b1 = warp_tot[2] + warp_tot[3];
        double b2 = warp_tot[4] + warp_tot[5], b3 = warp_tot[6] + warp_tot[7];
        double *l9 = t.W + level_off(t, lv + 9) + cta * 4;
        l9[0] = b0; l9[1] = b1; l9[2] = b2; l9[3] = b3;
        double c0 = b0 + b1, c1 = b2 + b3;
        double *l10 = t.W + level_off(t, lv + 10) + cta * 2;
        l10[0] = c0; l10[1] = c1;
        t.W[level_off(t, lv + 11) + cta] = c0 + c1;
        if (lv == 0) {
            double mx = warp_mx[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) mx = fmax(mx, warp_mx[k]);
            atomic_max_f64(&t.st->max_leaf, mx);
        }
    }
}

// Remaining levels (fewer than 2048 source nodes): one CTA, level by level.
__global__ void __launch_bounds__(1024) tree_build_top(TreeDev t, int lv, const double *src_leaves) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    if (lv == 0) {
        double mx = -INFINITY;
        for (int64_t i = tid; i < t.size; i += blockDim.x) {
            double x = src_leaves ? __ldg(src_leaves + i) : __ldcg(t.W + i);
            if (src_leaves) t.W[i] = x;
            mx = fmax(mx, x);
        }
        mx = mb::warp_max(mx);
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, red[k]);
            atomic_max_f64(&t.st->max_leaf, mx);
        }
    }
    for (int l = lv + 1; l < t.depth; ++l) {
        __syncthreads();
        const int64_t n = t.total >> (l + 1);
        const double *below = t.W + level_off(t, l - 1);
        double *here = t.W + level_off(t, l);
        for (int64_t j = tid; j < n; j += blockDim.x) here[j] = __ldcg(below + 2 * j) + __ldcg(below + 2 * j + 1);
    }
}

// ---------------------------------------------------------------------------------------------
// batched update  (update_leaf_batch, prioritized_buffer.py:152-186; _normalize_priority :430-434)
// ---------------------------------------------------------------------------------------------
// ---- large batches: the same algorithm as separate launches, sparse at the bottom, dense at the top ----
// The cooperative kernel above makes every entry recompute its ancestor on EVERY level: above the level
// where the tree has fewer nodes than the batch has entries, a million threads hammer a handful of
// addresses (measured 5.8 ms for 1 M updates of a 10 M-leaf tree, 0.016 of the HBM roofline).  Here the
// per-entry recompute only runs for the sparse bottom levels (nodes(l) > 2n); from there up the whole
// level is rebuilt with the dense build passes (a few MB).  Kernel boundaries replace the grid barriers.
template <typename WT>
__global__ void __launch_bounds__(256) tree_upd_stamp(TreeDev t, const int64_t *__restrict__ idx,
                                                      const uint8_t *__restrict__ valid, int64_t n, int gen) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (valid && !valid[i])) return;
    const int64_t k = idx[i];
    if (k < 0 || k >= t.size) {
        t.st->err = MB_ERR_INDEX_OOB;
        t.st->err_gen = gen;
    } else {
        atomicMax(t.owner + k, (int)i);
    }
}

template <typename WT>
__global__ void __launch_bounds__(256) tree_upd_leaves(TreeDev t, const int64_t *__restrict__ idx,
                                                       const WT *__restrict__ w, const uint8_t *__restrict__ valid,
                                                       int64_t n, Normalize nz, int gen) {
    if (*((volatile int *)&t.st->err_gen) == gen) return;  // rejected batch: nothing is written
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mx = -INFINITY;
    if (i < n && !(valid && !valid[i])) {
        const int64_t k = idx[i];
        const double x = normalized_weight(w[i], nz);
        mx = x;
        if (__ldcg(t.owner + k) == (int)i) t.W[k] = x;
    }
    mx = mb::warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomic_max_f64(&t.st->batch_max, mx);
}

__global__ void __launch_bounds__(256) tree_upd_level(TreeDev t, const int64_t *__restrict__ idx,
                                                      const uint8_t *__restrict__ valid, int64_t n, int l, int gen) {
    if (*((volatile int *)&t.st->err_gen) == gen) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (valid && !valid[i])) return;
    const int64_t node = idx[i] >> l;
    const double2 c = __ldcg(reinterpret_cast<const double2 *>(t.W + level_off(t, l - 1) + 2 * node));
    t.W[level_off(t, l) + node] = c.x + c.y;
}

__global__ void __launch_bounds__(256) tree_upd_finish(TreeDev t, const int64_t *__restrict__ idx,
                                                       const uint8_t *__restrict__ valid, int64_t n, int is_f32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const double bm = __ldcg(&t.st->batch_max);
        if (bm > -INFINITY && !(t.st->max_leaf > bm)) {
            t.st->max_leaf = bm;
            t.st->max_is_f32 = is_f32;
        }
        t.st->batch_max = -INFINITY;
    }
    if (i >= n || (valid && !valid[i])) return;
    const int64_t k = idx[i];
    if (k >= 0 && k < t.size) t.owner[k] = -1;
}

// update_leaf (:127-150): one path to the root.  The reference's bound check admits
// index == size (:135); that slot is a padding leaf whenever size is not a power of two.
__global__ void tree_update_one_kernel(TreeDev t, double weight, int64_t index) {
    if (threadIdx.x != 0) return;
    if (!(t.st->max_leaf > weight)) {  // max(weight, max_leaf): a Python float wins ties (:137)
        t.st->max_leaf = weight;
        t.st->max_is_f32 = 0;
    }
    t.W[index] = weight;
    int64_t node = index;
    for (int l = 1; l < t.depth; ++l) {
        node >>= 1;
        const double2 c = __ldcg(reinterpret_cast<const double2 *>(t.W + level_off(t, l - 1) + 2 * node));
        t.W[level_off(t, l) + node] = c.x + c.y;
        __threadfence_block();
    }
}

int ceil_log2_i64(int64_t x) {
    int l = 0;
    while (((int64_t)1 << l) < x) ++l;
    return l;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int mb_tree_create(int64_t size, int device, mb_tree **out) {
    if (!out || size < 1 || size > ((int64_t)1 << 31)) return MB_ERR_INVALID;
    if (mb_device_count() <= device || device < 0) return MB_ERR_NO_DEVICE;
    mb::DeviceGuard g(device);
    if (!g.ok) return MB_ERR_CUDA;
    mb_tree *t = new (std::nothrow) mb_tree();
    if (!t) return MB_ERR_NOMEM;
    t->device = device;
    const int depth = ceil_log2_i64(size) + 1;  // int(ceil(log2(size))) + 1 (:48)
    t->dev.size = size;
    t->dev.depth = depth;
    t->dev.total = (int64_t)1 << depth;
    t->n_leaf_pad = (int64_t)1 << (depth - 1);
    t->n_nodes = t->dev.total - 1;
    t->h_state = nullptr;
    t->gen = 0;
    {
        const char *e = getenv("MB_FORCE_GENERIC_UPDATE");
        t->force_generic = (e && e[0] == '1') ? 1 : 0;
        const char *lg = getenv("MB_UPDATE_LEGACY");
        t->legacy_update = (lg && lg[0] == '1') ? 1 : 0;
        const char *c = getenv("MB_SORTED_CLUSTER");
        t->sorted_cluster = (c && c[0] == '1') ? 1 : 0;
        const char *y = getenv("MB_SORTED_ASYNC");
        t->sorted_async = (y && y[0] == '1') ? 1 : 0;
        if (t->sorted_cluster || t->sorted_async) t->legacy_update = 1;   // asking for a legacy variant selects the legacy path
    }
    t->dev.W = nullptr;
    t->dev.owner = nullptr;
    t->dev.st = nullptr;
    if (cudaMalloc((void **)&t->dev.W, sizeof(double) * (size_t)t->dev.total) != cudaSuccess ||
        cudaMalloc((void **)&t->dev.owner, sizeof(int) * (size_t)size) != cudaSuccess ||
        cudaMalloc((void **)&t->dev.st, sizeof(TreeState)) != cudaSuccess ||
        cudaMallocHost((void **)&t->h_state, sizeof(TreeState)) != cudaSuccess) {
        cudaGetLastError();
        mb_tree_destroy(t);
        return MB_ERR_NOMEM;
    }
    MB_CUDA(cudaMemset(t->dev.W, 0, sizeof(double) * (size_t)t->dev.total));
    MB_CUDA(cudaMemset(t->dev.owner, 0xff, sizeof(int) * (size_t)size));
    MB_CUDA(cudaMemset(t->dev.st, 0, sizeof(TreeState)));
    tree_state_reset<<<1, 1>>>(t->dev.st, 0.0);
    MB_CUDA(cudaFuncSetAttribute(tree_update_sorted<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    MB_CUDA(cudaFuncSetAttribute(tree_update_sorted<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    MB_CUDA(cudaFuncSetAttribute(tree_update_part<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    MB_CUDA(cudaFuncSetAttribute(tree_update_part<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    {   // opt-in variants (MB_SORTED_CLUSTER / MB_SORTED_ASYNC): a failure here must not take the default path down
        const void *v2[] = {(const void *)tree_update_sorted_v2<float, 3, false>, (const void *)tree_update_sorted_v2<double, 3, false>,
                            (const void *)tree_update_sorted_v2<float, 3, true>, (const void *)tree_update_sorted_v2<double, 3, true>,
                            (const void *)tree_update_sorted_v2<float, 0, true>, (const void *)tree_update_sorted_v2<double, 0, true>};
        for (const void *f : v2)
            if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) {
                cudaGetLastError();
                t->sorted_cluster = t->sorted_async = 0;
            }
    }
    int per_sm = 0;
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_update_kernel<double, true>, 1024, 0));
    int sms = mb::kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    t->coop_blocks = (per_sm > 0 ? per_sm : 1) * sms;
    MB_CUDA(cudaDeviceSynchronize());
    *out = t;
    return MB_OK;
}

int mb_tree_destroy(mb_tree *t) {
    if (!t) return MB_OK;
    mb::DeviceGuard g(t->device);
    if (t->dev.W) cudaFree(t->dev.W);
    if (t->dev.owner) cudaFree(t->dev.owner);
    if (t->dev.st) cudaFree(t->dev.st);
    if (t->h_state) cudaFreeHost(t->h_state);
    delete t;
    return MB_OK;
}

int mb_tree_geometry(const mb_tree *t, int64_t *size, int *depth, int64_t *n_leaf_padded, int64_t *n_nodes) {
    if (!t) return MB_ERR_INVALID;
    if (size) *size = t->dev.size;
    if (depth) *depth = t->dev.depth;
    if (n_leaf_padded) *n_leaf_padded = t->n_leaf_pad;
    if (n_nodes) *n_nodes = t->n_nodes;
    return MB_OK;
}

double *mb_tree_weights(mb_tree *t) { return t ? t->dev.W : nullptr; }
const double *mb_tree_root_ptr(mb_tree *t) { return t ? t->dev.W + t->n_nodes - 1 : nullptr; }
const double *mb_tree_max_leaf_ptr(mb_tree *t) { return t ? &t->dev.st->max_leaf : nullptr; }
void *mb_tree_state_blob(mb_tree *t, int64_t *nbytes) {
    if (nbytes) *nbytes = (int64_t)sizeof(TreeState);
    return t ? (void *)t->dev.st : nullptr;
}

int mb_tree_stats(mb_tree *t, double *h_weight_sum, double *h_max_leaf, int *h_max_leaf_dtype, void *stream) {
    if (!t) return MB_ERR_INVALID;
    mb::DeviceGuard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    MB_CUDA(cudaMemcpyAsync(&t->h_state->isw_max, t->dev.W + t->n_nodes - 1, sizeof(double),
                            cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaMemcpyAsync(&t->h_state->max_leaf, &t->dev.st->max_leaf, sizeof(double),
                            cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaMemcpyAsync(&t->h_state->max_is_f32, &t->dev.st->max_is_f32, sizeof(int),
                            cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    if (h_weight_sum) *h_weight_sum = t->h_state->isw_max;
    if (h_max_leaf) *h_max_leaf = t->h_state->max_leaf;
    if (h_max_leaf_dtype) *h_max_leaf_dtype = t->h_state->max_is_f32 ? MB_F32 : MB_F64;
    return MB_OK;
}

int mb_tree_check(mb_tree *t, void *stream) {
    if (!t) return MB_ERR_INVALID;
    mb::DeviceGuard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    MB_CUDA(cudaMemcpyAsync(&t->h_state->err, &t->dev.st->err, sizeof(int), cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    const int err = t->h_state->err;
    if (err != MB_OK) {
        MB_CUDA(cudaMemsetAsync(&t->dev.st->err, 0, sizeof(int), s));
        MB_CUDA(cudaStreamSynchronize(s));
    }
    return err;
}

int mb_tree_clear(mb_tree *t, void *stream) {
    if (!t) return MB_ERR_INVALID;
    mb::DeviceGuard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    MB_CUDA(cudaMemsetAsync(t->dev.W, 0, sizeof(double) * (size_t)t->dev.total, s));
    MB_CUDA(cudaMemsetAsync(t->dev.st, 0, sizeof(TreeState), s));
    tree_state_reset<<<1, 1, 0, s>>>(t->dev.st, 0.0);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

int mb_tree_build(mb_tree *t, const double *d_leaves, int64_t n, void *stream) {
    if (!t) return MB_ERR_INVALID;
    if (d_leaves && n != t->dev.size) return MB_ERR_SIZE_MISMATCH;
    mb::DeviceGuard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    tree_state_reset<<<1, 1, 0, s>>>(t->dev.st, -INFINITY);  // _build replaces max_leaf (:220)
    int lv = 0;
    while (t->dev.depth - 1 - lv >= 11) {
        const int64_t n_src = t->dev.total >> (lv + 1);
        tree_build_pass<<<(unsigned)(n_src / 2048), 256, 0, s>>>(t->dev, lv, lv == 0 ? d_leaves : nullptr);
        lv += 11;
    }
    if (lv < t->dev.depth - 1 || lv == 0)
        tree_build_top<<<1, 1024, 0, s>>>(t->dev, lv, lv == 0 ? d_leaves : nullptr);
    MB_LAUNCHED(2 + (t->dev.depth - 1) / 11);
    return MB_OK;
}

int mb_tree_find(mb_tree *t, const double *d_query, int64_t n, int64_t *d_index, void *stream) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    mb::DeviceGuard g(t->device);
    tree_find_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t->dev, d_query, n, d_index);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_tree_leaf_weight(mb_tree *t, const int64_t *d_index, int64_t n, double *d_weight, void *stream) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    mb::DeviceGuard g(t->device);
    tree_leaf_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t->dev, d_index, n, d_weight);
    MB_LAUNCHED(1);
    return MB_OK;
}

static int launch_update(mb_tree *t, const int64_t *d_index, const void *d_weight, int dtype,
                         const uint8_t *d_valid, int64_t n, Normalize nz, cudaStream_t s) {
    if (!t || n < 0 || (dtype != MB_F32 && dtype != MB_F64)) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    if (n > 0x7fffffff) return MB_ERR_INVALID;
    mb::DeviceGuard g(t->device);
    const int gen = ++t->gen;
    if (n <= kPartMaxN && !t->force_generic && !t->legacy_update) {
        // one launch, one CTA per subtree (128 at 10 M leaves), any order / mask / duplicates
        const int levels = t->dev.depth - 1;
        const unsigned grid = 1u << part_cut(levels);
        const size_t dyn = part_dyn_smem(levels);
        if (dtype == MB_F32) {
            MB_PREFER_MAX_SMEM(tree_update_part<float>);
            tree_update_part<float><<<grid, kPartThreads, dyn, s>>>(t->dev, d_index, (const float *)d_weight, d_valid, (int)n, nz, gen);
        } else {
            MB_PREFER_MAX_SMEM(tree_update_part<double>);
            tree_update_part<double><<<grid, kPartThreads, dyn, s>>>(t->dev, d_index, (const double *)d_weight, d_valid, (int)n, nz, gen);
        }
        MB_LAUNCHED(1);
        return MB_OK;
    }
    if (n <= kSortedN && !t->force_generic) {
        const int threads = (int)(((n + 31) / 32) * 32);
        const size_t dyn = sizeof(double) * (size_t)kSortedN * (size_t)(t->dev.depth - 1);
        const bool cluster = t->sorted_cluster && n > 64 && t->dev.depth - 1 >= 3 + 4;
        if (cluster || t->sorted_async) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cluster ? 8 : 1);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = dyn;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cluster ? 8 : 1;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = cluster ? 1 : 0;
            const int ni = (int)n;
            const float *wf = (const float *)d_weight;
            const double *wd = (const double *)d_weight;
            cudaError_t e;
            if (cluster && t->sorted_async)
                e = dtype == MB_F32 ? cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<float, 3, true>, t->dev, d_index, wf, d_valid, ni, nz, gen)
                                    : cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<double, 3, true>, t->dev, d_index, wd, d_valid, ni, nz, gen);
            else if (cluster)
                e = dtype == MB_F32 ? cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<float, 3, false>, t->dev, d_index, wf, d_valid, ni, nz, gen)
                                    : cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<double, 3, false>, t->dev, d_index, wd, d_valid, ni, nz, gen);
            else
                e = dtype == MB_F32 ? cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<float, 0, true>, t->dev, d_index, wf, d_valid, ni, nz, gen)
                                    : cudaLaunchKernelEx(&cfg, tree_update_sorted_v2<double, 0, true>, t->dev, d_index, wd, d_valid, ni, nz, gen);
            MB_CUDA(e);
            mb::count_launch(1);
            return MB_OK;
        }
        if (dtype == MB_F32)
            tree_update_sorted<float><<<1, threads, dyn, s>>>(t->dev, d_index, (const float *)d_weight, d_valid, (int)n, nz, gen);
        else
            tree_update_sorted<double><<<1, threads, dyn, s>>>(t->dev, d_index, (const double *)d_weight, d_valid, (int)n, nz, gen);
        MB_LAUNCHED(1);
        return MB_OK;
    }
    if (n <= 1024 || (t->force_generic && n <= 4096)) {
        const int threads = (int)(n <= 1024 ? ((n + 31) / 32) * 32 : 1024);
        if (dtype == MB_F32)
            tree_update_kernel<float, false><<<1, threads, 0, s>>>(t->dev, d_index, (const float *)d_weight, d_valid, n, nz, gen);
        else
            tree_update_kernel<double, false><<<1, threads, 0, s>>>(t->dev, d_index, (const double *)d_weight, d_valid, n, nz, gen);
        MB_LAUNCHED(1);
        return MB_OK;
    }
    if (!t->force_generic) {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        const int depth = t->dev.depth;
        int ls = 1;  // first level that is rebuilt densely: nodes(ls) <= 2n
        while (ls < depth - 1 && (t->dev.total >> (ls + 1)) > 2 * n) ++ls;
        if (dtype == MB_F32) {
            tree_upd_stamp<float><<<blocks, 256, 0, s>>>(t->dev, d_index, d_valid, n, gen);
            tree_upd_leaves<float><<<blocks, 256, 0, s>>>(t->dev, d_index, (const float *)d_weight, d_valid, n, nz, gen);
        } else {
            tree_upd_stamp<double><<<blocks, 256, 0, s>>>(t->dev, d_index, d_valid, n, gen);
            tree_upd_leaves<double><<<blocks, 256, 0, s>>>(t->dev, d_index, (const double *)d_weight, d_valid, n, nz, gen);
        }
        int launches = 3;
        for (int l = 1; l <= ls; ++l, ++launches) tree_upd_level<<<blocks, 256, 0, s>>>(t->dev, d_index, d_valid, n, l, gen);
        int lv = ls;
        for (; depth - 1 - lv >= 11; lv += 11, ++launches)
            tree_build_pass<<<(unsigned)((t->dev.total >> (lv + 1)) / 2048), 256, 0, s>>>(t->dev, lv, nullptr);
        if (lv < depth - 1) {
            tree_build_top<<<1, 1024, 0, s>>>(t->dev, lv, nullptr);
            ++launches;
        }
        tree_upd_finish<<<blocks, 256, 0, s>>>(t->dev, d_index, d_valid, n, dtype == MB_F32 ? 1 : 0);
        MB_LAUNCHED(launches);
        return MB_OK;
    }
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > t->coop_blocks) blocks = t->coop_blocks;
    void *args[] = {&t->dev, &d_index, &d_weight, &d_valid, &n, &nz, (void *)&gen};
    const void *fn = dtype == MB_F32 ? (const void *)tree_update_kernel<float, true>
                                     : (const void *)tree_update_kernel<double, true>;
    MB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(1024), args, 0, s));
    mb::count_launch(1);
    return MB_OK;
}

int mb_tree_update_batch(mb_tree *t, const int64_t *d_index, const void *d_weight, int weight_dtype,
                         const uint8_t *d_valid, int64_t n, void *stream) {
    Normalize nz{0, 0.0, 0.0};
    return launch_update(t, d_index, d_weight, weight_dtype, d_valid, n, nz, (cudaStream_t)stream);
}

int mb_tree_update_priority(mb_tree *t, const int64_t *d_index, const void *d_priority, int priority_dtype,
                            const uint8_t *d_valid, int64_t n, double epsilon, double alpha, void *stream) {
    Normalize nz{1, epsilon, alpha};
    return launch_update(t, d_index, d_priority, priority_dtype, d_valid, n, nz, (cudaStream_t)stream);
}

int mb_tree_update_leaf(mb_tree *t, double weight, int64_t index, void *stream) {
    if (!t) return MB_ERR_INVALID;
    // reference check is `0 <= index <= size` (:135); index == size only exists as a padding leaf
    if (index < 0 || index > t->dev.size || index >= t->n_leaf_pad) return MB_ERR_INDEX_OOB;
    mb::DeviceGuard g(t->device);
    tree_update_one_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t->dev, weight, index);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_per_sample(mb_tree *t, const double *d_uniform, int64_t n, int64_t n_stored, double all_weight_sum,
                  double beta, const mb_per_params *d_params, int64_t *d_index, double *d_is_weight,
                  void *stream) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    mb::DeviceGuard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    PerArgs a{n_stored, all_weight_sum, beta, d_params};
    if (n <= 1024 && n > 64 && !t->force_generic) {
        const unsigned threads = n <= 256 ? 32 : (n <= 512 ? 64 : 128);
        const unsigned ctas = (unsigned)((n + threads - 1) / threads);  // <= 8: one cluster
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(threads);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ctas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MB_PREFER_MAX_SMEM(per_sample_cluster);
        MB_CUDA(cudaLaunchKernelEx(&cfg, per_sample_cluster, t->dev, d_uniform, n, a, d_index, d_is_weight));
    } else if (n <= 1024) {
        per_sample_small<<<1, (unsigned)(((n + 31) / 32) * 32), 0, s>>>(t->dev, d_uniform, n, a, d_index, d_is_weight);
    } else {
        per_isw_reset<<<1, 1, 0, s>>>(t->dev.st);
        per_sample_large<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t->dev, d_uniform, n, a, d_index, d_is_weight);
        per_normalize<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t->dev.st, d_is_weight, n);
        mb::count_launch(2);
    }
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_per_sample_multi(mb_tree *t, const double *d_uniform, int64_t n, int n_requesters,
                        const mb_per_params *d_params, int64_t *d_index, double *d_is_weight, void *stream) {
    if (!t || n < 0 || n_requesters < 1 || !d_params || !d_uniform || !d_index || !d_is_weight) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    if (!(n <= 1024 && n > 64) || t->force_generic) {      // one launch per requester through the ordinary entry point
        for (int r = 0; r < n_requesters; ++r) {
            const int rc = mb_per_sample(t, d_uniform + (int64_t)r * n, n, 0, 0.0, 0.0, d_params + r, d_index + (int64_t)r * n,
                                         d_is_weight + (int64_t)r * n, stream);
            if (rc != MB_OK) return rc;
        }
        return MB_OK;
    }
    mb::DeviceGuard g(t->device);
    const unsigned threads = n <= 256 ? 32 : (n <= 512 ? 64 : 128);
    const unsigned ctas = (unsigned)((n + threads - 1) / threads);  // <= 8: one cluster per requester
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas * (unsigned)n_requesters);
    cfg.blockDim = dim3(threads);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MB_PREFER_MAX_SMEM(per_sample_cluster_multi);
    MB_CUDA(cudaLaunchKernelEx(&cfg, per_sample_cluster_multi, t->dev, d_uniform, n, d_params, d_index, d_is_weight));
    MB_LAUNCHED(1);
    return MB_OK;
}

// ------------------------------------------------------------------------------ host variants
#define MB_STAGE(t, hb, db)                                   \
    mb::DeviceGuard g((t)->device);                           \
    mb::Staging *st = mb::staging_for((t)->device);           \
    if (!st) return MB_ERR_CUDA;                              \
    std::lock_guard<std::mutex> _staging_lock(st->mu);        \
    {                                                         \
        int _r = st->ensure((hb), (db));                      \
        if (_r != MB_OK) return _r;                           \
    }

int mb_tree_build_host(mb_tree *t, const double *h_leaves, int64_t n) {
    if (!t || !h_leaves) return MB_ERR_INVALID;
    if (n != t->dev.size) return MB_ERR_SIZE_MISMATCH;
    const size_t bytes = sizeof(double) * (size_t)n;
    MB_STAGE(t, bytes, bytes);
    memcpy(st->h, h_leaves, bytes);
    MB_CUDA(cudaMemcpyAsync(st->d, st->h, bytes, cudaMemcpyHostToDevice, st->stream));
    int r = mb_tree_build(t, (const double *)st->d, n, st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaStreamSynchronize(st->stream));
    return MB_OK;
}

int mb_tree_find_host(mb_tree *t, const double *h_query, int64_t n, int64_t *h_index) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    const size_t bytes = 8 * (size_t)n;
    MB_STAGE(t, 2 * bytes, 2 * bytes);
    memcpy(st->h, h_query, bytes);
    MB_CUDA(cudaMemcpyAsync(st->d, st->h, bytes, cudaMemcpyHostToDevice, st->stream));
    int r = mb_tree_find(t, (const double *)st->d, n, (int64_t *)(st->d + bytes), st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaMemcpyAsync(st->h + bytes, st->d + bytes, bytes, cudaMemcpyDeviceToHost, st->stream));
    MB_CUDA(cudaStreamSynchronize(st->stream));
    memcpy(h_index, st->h + bytes, bytes);
    return MB_OK;
}

int mb_tree_leaf_weight_host(mb_tree *t, const int64_t *h_index, int64_t n, double *h_weight) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    for (int64_t i = 0; i < n; ++i)
        if (h_index[i] < 0 || h_index[i] >= t->dev.size) return MB_ERR_INDEX_OOB;  // (:89-90)
    const size_t bytes = 8 * (size_t)n;
    MB_STAGE(t, 2 * bytes, 2 * bytes);
    memcpy(st->h, h_index, bytes);
    MB_CUDA(cudaMemcpyAsync(st->d, st->h, bytes, cudaMemcpyHostToDevice, st->stream));
    int r = mb_tree_leaf_weight(t, (const int64_t *)st->d, n, (double *)(st->d + bytes), st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaMemcpyAsync(st->h + bytes, st->d + bytes, bytes, cudaMemcpyDeviceToHost, st->stream));
    MB_CUDA(cudaStreamSynchronize(st->stream));
    memcpy(h_weight, st->h + bytes, bytes);
    return MB_OK;
}

int mb_tree_update_batch_host(mb_tree *t, const int64_t *h_index, const void *h_weight, int weight_dtype,
                              int64_t n) {
    if (!t || n < 0 || (weight_dtype != MB_F32 && weight_dtype != MB_F64)) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    for (int64_t i = 0; i < n; ++i)
        if (h_index[i] < 0 || h_index[i] >= t->dev.size) return MB_ERR_INDEX_OOB;  // (:172-173)
    const size_t ib = 8 * (size_t)n, wb = (weight_dtype == MB_F32 ? 4 : 8) * (size_t)n;
    MB_STAGE(t, ib + wb, ib + wb);
    memcpy(st->h, h_index, ib);
    memcpy(st->h + ib, h_weight, wb);
    MB_CUDA(cudaMemcpyAsync(st->d, st->h, ib + wb, cudaMemcpyHostToDevice, st->stream));
    int r = mb_tree_update_batch(t, (const int64_t *)st->d, st->d + ib, weight_dtype, nullptr, n, st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaStreamSynchronize(st->stream));
    return MB_OK;
}

int mb_tree_read_host(mb_tree *t, int64_t first_node, int64_t n_nodes, double *h_out) {
    if (!t || first_node < 0 || n_nodes < 0 || first_node + n_nodes > t->n_nodes) return MB_ERR_INVALID;
    if (n_nodes == 0) return MB_OK;
    mb::DeviceGuard g(t->device);
    MB_CUDA(cudaMemcpy(h_out, t->dev.W + first_node, sizeof(double) * (size_t)n_nodes, cudaMemcpyDeviceToHost));
    return MB_OK;
}

int mb_per_sample_host(mb_tree *t, const double *h_uniform, int64_t n, int64_t n_stored, double all_weight_sum,
                       double beta, int64_t *h_index, double *h_is_weight) {
    if (!t || n < 0) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    const size_t bytes = 8 * (size_t)n;
    MB_STAGE(t, 3 * bytes, 3 * bytes);
    memcpy(st->h, h_uniform, bytes);
    MB_CUDA(cudaMemcpyAsync(st->d, st->h, bytes, cudaMemcpyHostToDevice, st->stream));
    int r = mb_per_sample(t, (const double *)st->d, n, n_stored, all_weight_sum, beta, nullptr,
                          (int64_t *)(st->d + bytes), (double *)(st->d + 2 * bytes), st->stream);
    if (r != MB_OK) return r;
    MB_CUDA(cudaMemcpyAsync(st->h + bytes, st->d + bytes, 2 * bytes, cudaMemcpyDeviceToHost, st->stream));
    MB_CUDA(cudaStreamSynchronize(st->stream));
    memcpy(h_index, st->h + bytes, bytes);
    memcpy(h_is_weight, st->h + 2 * bytes, bytes);
    return MB_OK;
}

}  // extern "C"
