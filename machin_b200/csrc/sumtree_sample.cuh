// Sum-tree descent and the fused PER draw: device code only (no CUDA runtime calls).  Included by sumtree.cu
// inside its anonymous namespace (after sumtree_update.cuh, which provides level_off) and by
// tests/cpu_emul/sample_harness.cpp, which compiles THIS SOURCE with g++ against the CUDA shim.
#pragma once
#include "sumtree_types.cuh"

__device__ __forceinline__ double2 ld_nc2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// ---------------------------------------------------------------------------------------------
// descent  (find_leaf_index, prioritized_buffer.py:96-125)
// ---------------------------------------------------------------------------------------------
// Descend J levels from node `idx` whose 2^J descendants start at W[child_off + idx * 2^J].
template <int J>
__device__ __forceinline__ void descend(const double *__restrict__ W, int64_t child_off, int64_t &idx, double &w) {
    constexpr int N = 1 << J;
    double tr[2 * N];  // heap order: tr[1] = current node, tr[N + k] = k-th descendant
    const double *p = W + child_off + idx * N;
#pragma unroll
    for (int k = 0; k < N; k += 2) {
        double2 a = ld_nc2(p + k);
        tr[N + k] = a.x;
        tr[N + k + 1] = a.y;
    }
#pragma unroll
    for (int i = N - 1; i >= 2; --i) tr[i] = tr[2 * i] + tr[2 * i + 1];
    int rel = 0;  // position inside the current level of the mini tree
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double left = 0.0;
#pragma unroll
        for (int m = 0; m < (1 << j); ++m)
            if (m == rel) left = tr[2 * ((1 << j) + m)];
        const bool right = w > left;  // strict, like `weight > left_wt` (:117)
        if (right) w = w - left;      // `weight - left_wt * select` (:119)
        rel = 2 * rel + (right ? 1 : 0);
    }
    idx = idx * N + rel;
}

__device__ __forceinline__ int64_t tree_find_one(const TreeDev &t, double w) {
    int64_t idx = 0;
    int remaining = t.depth - 1;  // levels to descend
    const int head = remaining & 3;
    if (head == 1) descend<1>(t.W, level_off(t, remaining - 1), idx, w);
    else if (head == 2) descend<2>(t.W, level_off(t, remaining - 2), idx, w);
    else if (head == 3) descend<3>(t.W, level_off(t, remaining - 3), idx, w);
    remaining -= head;
    while (remaining > 0) {
        remaining -= 4;
        descend<4>(t.W, level_off(t, remaining), idx, w);
    }
    return min(max(idx, (int64_t)0), t.size - 1);  // np.clip(index, 0, size - 1) (:120)
}

__global__ void __launch_bounds__(256) tree_find_kernel(TreeDev t, const double *__restrict__ q, int64_t n,
                                                        int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tree_find_one(t, q[i]);
}

__global__ void tree_leaf_weight_kernel(TreeDev t, const int64_t *__restrict__ idx, int64_t n,
                                        double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = idx[i];
    if (k < 0 || k >= t.size) {
        t.st->err = MB_ERR_INDEX_OOB;
        out[i] = 0.0;
    } else {
        out[i] = __ldcg(t.W + k);
    }
}

// ---------------------------------------------------------------------------------------------
// fused PER draw  (sample_index_and_weight, prioritized_buffer.py:397-428)
// ---------------------------------------------------------------------------------------------
struct PerArgs {
    int64_t n_stored;
    double all_weight_sum;
    double beta;
    const mb_per_params *d_params;
};

__device__ __forceinline__ void per_one(const TreeDev &t, const double *__restrict__ u, int64_t n, int64_t i,
                                        const PerArgs &a, int64_t &index, double &isw) {
    double beta = a.beta, total = a.all_weight_sum;
    int64_t n_stored = a.n_stored, row = 0;
    if (a.d_params) {
        beta = a.d_params->beta;
        total = a.d_params->all_weight_sum;
        n_stored = a.d_params->n_stored;
        row = a.d_params->step;
    }
    // (n is the number of draws actually made: callers pass per_draws(a, launch n))
    const double root = __ldcg(t.W + t.total - 2);
    // three separately rounded float64 ops, no FMA contraction (:409-412)
    const double seg = __ddiv_rn(root, (double)n);
    double q = __dadd_rn(__dmul_rn(u[row * n + i], seg), __dmul_rn((double)i, seg));
    const double hi = (root - 1e-6 > 0.0) ? root - 1e-6 : 0.0;  // max(sum - 1e-6, 0) (:414)
    q = fmin(fmax(q, 0.0), hi);                                  // np.clip (:413)
    index = tree_find_one(t, q);
    const double p = __ldg(t.W + index);
    if (!(total > 0.0) && !(total < 0.0)) total = root;          // `all_weight_sum or root` (:421)
    const double prob = __ddiv_rn(p, total);
    isw = pow(__dmul_rn((double)n_stored, prob), -beta);         // (:423)
}

// number of draws: the launch's n, or a device-side override (fixed-shape graph replay, see mb_per_params)
__device__ __forceinline__ int64_t per_draws(const PerArgs &a, int64_t n) {
    if (a.d_params) {
        const int64_t d = a.d_params->n_draw;
        if (d > 0 && d < n) return d;
    }
    return n;
}

// n <= 1024: one CTA does the draw and the max-normalisation (:424).
__global__ void __launch_bounds__(1024) per_sample_small(TreeDev t, const double *__restrict__ u, int64_t n,
                                                         PerArgs a, int64_t *__restrict__ out_idx,
                                                         double *__restrict__ out_isw) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    n = per_draws(a, n);
    int64_t index = 0;
    double isw = -INFINITY;
    if (tid < n) per_one(t, u, n, tid, a, index, isw);
    double mx = mb::warp_max(isw);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < (int)((blockDim.x + 31) >> 5); ++k) mx = fmax(mx, red[k]);
    if (tid < n) {
        out_idx[tid] = index;
        out_isw[tid] = __ddiv_rn(isw, mx);
    }
}

// n <= 1024, B200 path: the batch is spread over a thread-block CLUSTER of up to 8 CTAs (8 SMs), because
// a single SM's L1 serves ~1 wavefront per cycle and thread-per-query descent issues 32 distinct lines per
// load instruction (4096 wavefronts per jump for 512 queries on one SM).  The max-normalisation crosses
// CTAs through distributed shared memory: each CTA publishes its maximum, cluster.sync(), every CTA reads
// the others' maxima over DSMEM.
__global__ void __launch_bounds__(128) per_sample_cluster(TreeDev t, const double *__restrict__ u, int64_t n,
                                                          PerArgs a, int64_t *__restrict__ out_idx,
                                                          double *__restrict__ out_isw) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double red[4];
    __shared__ double cta_max;
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + tid;
    n = per_draws(a, n);
    int64_t index = 0;
    double isw = -INFINITY;
    if (i < n) per_one(t, u, n, i, a, index, isw);
    double mx = mb::warp_max(isw);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, red[k]);
        cta_max = mx;
    }
    cluster.sync();
    mx = -INFINITY;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) mx = fmax(mx, *cluster.map_shared_rank(&cta_max, r));
    cluster.sync();  // nobody exits (and frees its shared memory) while a peer may still read it
    if (i < n) {
        out_idx[i] = index;
        out_isw[i] = __ddiv_rn(isw, mx);
    }
}

// The same draw for R requesters in ONE launch (fixed-shape shard exchange: a shard serves every requester of an
// update): cluster c of the grid handles requester c -- its own uniforms (row c of u[R][n]), its own device-side
// parameter block (params[c]: beta, counts, n_draw) and its own output rows.  Clusters are independent.
__global__ void __launch_bounds__(128) per_sample_cluster_multi(TreeDev t, const double *__restrict__ u, int64_t n,
                                                                const mb_per_params *__restrict__ params,
                                                                int64_t *__restrict__ out_idx,
                                                                double *__restrict__ out_isw) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double red[4];
    __shared__ double cta_max;
    const int tid = threadIdx.x;
    const unsigned ctas = cluster.num_blocks();
    const int64_t req = blockIdx.x / ctas;
    const int64_t i = (int64_t)(blockIdx.x % ctas) * blockDim.x + tid;
    const PerArgs a{0, 0.0, 0.0, params + req};
    const int64_t draws = per_draws(a, n);
    u += req * n;                          // (per_one reads u[step * draws + i]; step is 0 in this mode)
    int64_t index = 0;
    double isw = -INFINITY;
    if (i < draws) per_one(t, u, draws, i, a, index, isw);
    double mx = mb::warp_max(isw);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, red[k]);
        cta_max = mx;
    }
    cluster.sync();
    mx = -INFINITY;
    for (unsigned r = 0; r < ctas; ++r) mx = fmax(mx, *cluster.map_shared_rank(&cta_max, r));
    cluster.sync();  // nobody exits (and frees its shared memory) while a peer may still read it
    if (i < draws) {
        out_idx[req * n + i] = index;
        out_isw[req * n + i] = __ddiv_rn(isw, mx);
    }
}

__global__ void __launch_bounds__(256) per_sample_large(TreeDev t, const double *__restrict__ u, int64_t n,
                                                        PerArgs a, int64_t *__restrict__ out_idx,
                                                        double *__restrict__ out_isw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t index = 0;
    double isw = -INFINITY;
    if (i < n) {
        per_one(t, u, n, i, a, index, isw);
        out_idx[i] = index;
        out_isw[i] = isw;
    }
    const double mx = mb::warp_max(isw);
    if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomic_max_f64(&t.st->isw_max, mx);
}

__global__ void per_isw_reset(TreeState *st) { st->isw_max = -INFINITY; }

__global__ void per_normalize(const TreeState *st, double *__restrict__ isw, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) isw[i] = __ddiv_rn(isw[i], st->isw_max);
}
