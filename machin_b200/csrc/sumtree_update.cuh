// Sum-tree update kernels for small batches: device code only (no CUDA runtime calls).  Included by sumtree.cu
// inside its anonymous namespace and by tests/cpu_emul/tree_harness.cpp, which compiles THIS SOURCE with g++
// against a shim of the CUDA execution model (threads, warps, CTA barriers, a cluster of CTAs with distributed
// shared memory) so that the kernels written while no GPU was available can be checked against the oracle in the
// build container.  Expects `cg` (cooperative_groups), `mb::warp_max` and machin_b200.h to be visible.
#pragma once
#include "sumtree_types.cuh"

__device__ __forceinline__ int64_t level_off(const TreeDev &t, int l) { return t.total - (t.total >> l); }

__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v > __longlong_as_double((long long)old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

struct Normalize {
    int enabled;
    double epsilon, alpha;
};

// (|p| + eps) ** alpha in the dtype of p.  float32: NumPy evaluates |p| + f32(eps) in float32 and
// powf(x, f32(alpha)); we evaluate pow in float64 on the float32-rounded operands and round once,
// which equals a correctly rounded powf except within ~2^-29 of a rounding boundary.
__device__ __forceinline__ double normalized_weight(float p, const Normalize &nz) {
    if (!nz.enabled) return (double)p;
    const float x = fabsf(p) + (float)nz.epsilon;
    return (double)(float)pow((double)x, (double)(float)nz.alpha);
}
__device__ __forceinline__ double normalized_weight(double p, const Normalize &nz) {
    if (!nz.enabled) return p;
    return pow(fabs(p) + nz.epsilon, nz.alpha);
}

template <bool kCoop>
__device__ __forceinline__ void phase_sync() {
    if (kCoop) cg::this_grid().sync();
    else __syncthreads();
}

// Phases (each separated by a CTA-wide or grid-wide barrier):
//   0  validate indexes, stamp owner[idx] = max(position)        -> last write wins
//   1  winners store their leaf; max_leaf = max(max_leaf, max(w)) over ALL entries (:175)
//   l  for every level l = 1..depth-1: parent(idx >> l) = left + right   (idempotent, so the
//      duplicates that np.unique removes in the reference are simply recomputed)
//   e  clear the stamps
template <typename WT, bool kCoop>
__device__ __forceinline__ void tree_update_body(const TreeDev &t, const int64_t *__restrict__ idx,
                                                 const WT *__restrict__ w, const uint8_t *__restrict__ valid,
                                                 int64_t n, const Normalize &nz, int gen) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += stride) {
        if (valid && !valid[i]) continue;
        const int64_t k = idx[i];
        if (k < 0 || k >= t.size) {
            t.st->err = MB_ERR_INDEX_OOB;
            t.st->err_gen = gen;
        } else {
            atomicMax(t.owner + k, (int)i);
        }
    }
    __threadfence();
    phase_sync<kCoop>();
    const bool bad = *((volatile int *)&t.st->err_gen) == gen;
    if (!bad) {
        double mx = -INFINITY;
        for (int64_t i = tid; i < n; i += stride) {
            if (valid && !valid[i]) continue;
            const int64_t k = idx[i];
            const double x = normalized_weight(w[i], nz);
            mx = fmax(mx, x);
            if (__ldcg(t.owner + k) == (int)i) t.W[k] = x;
        }
        mx = mb::warp_max(mx);
        if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomic_max_f64(&t.st->batch_max, mx);
        for (int l = 1; l < t.depth; ++l) {
            phase_sync<kCoop>();
            const double *below = t.W + level_off(t, l - 1);
            double *here = t.W + level_off(t, l);
            for (int64_t i = tid; i < n; i += stride) {
                if (valid && !valid[i]) continue;
                const int64_t node = idx[i] >> l;
                const double2 c = __ldcg(reinterpret_cast<const double2 *>(below + 2 * node));
                here[node] = c.x + c.y;
            }
        }
    }
    phase_sync<kCoop>();
    for (int64_t i = tid; i < n; i += stride) {
        if (valid && !valid[i]) continue;
        const int64_t k = idx[i];
        if (k >= 0 && k < t.size) t.owner[k] = -1;
    }
    if (tid == 0) {
        // max_leaf = max(np.max(weights), max_leaf): the new value (and its dtype) wins ties (:175)
        const double bm = __ldcg(&t.st->batch_max);
        if (bm > -INFINITY && !(t.st->max_leaf > bm)) {
            t.st->max_leaf = bm;
            t.st->max_is_f32 = sizeof(WT) == 4 ? 1 : 0;
        }
        t.st->batch_max = -INFINITY;
    }
}

template <typename WT, bool kCoop>
__global__ void __launch_bounds__(1024) tree_update_kernel(TreeDev t, const int64_t *__restrict__ idx,
                                                           const WT *__restrict__ w,
                                                           const uint8_t *__restrict__ valid, int64_t n,
                                                           Normalize nz, int gen) {
    tree_update_body<WT, kCoop>(t, idx, w, valid, n, nz, gen);
}

// ---- small SORTED batches (n <= 512): the learner's priority write-back -------------------------------
// Indexes that come out of the stratified PER draw are ascending.  For a sorted batch the set of touched
// nodes on every level, who merges with whom and on which level, is known up front from neighbouring keys:
// entry i (after dropping all but the last of equal keys) represents its ancestor up to level
// hb(i) = highest bit in which its key differs from its predecessor's; an entry whose hb equals the
// current level hands its node value to the entry on its left and retires.  So the walk needs no hash
// table and no atomics: per level one barrier, one shared-memory read, one add.  The stored sibling of
// every node a thread will represent is prefetched up front (all loads in flight at once), results are
// parked in shared memory and written out after the walk (a barrier after a global store waits for L2).
// Parents are fl(left + right) of the current children: bit-identical to the level-synchronous kernel.
// Unsorted input (or a validity mask) takes the generic body in the same launch.
constexpr int kSortedN = 512;
constexpr int kSortedLevels = 31;

template <typename WT>
__global__ void __launch_bounds__(kSortedN) tree_update_sorted(TreeDev t, const int64_t *__restrict__ idx,
                                                               const WT *__restrict__ w,
                                                               const uint8_t *__restrict__ valid, int n,
                                                               Normalize nz, int gen) {
    MB_DYNAMIC_SMEM(dyn);
    double *park = reinterpret_cast<double *>(dyn);   // [levels][kSortedN]: stored siblings, then new node values
    __shared__ double s_val[kSortedN];                // value handed to the left neighbour
    __shared__ unsigned s_key[kSortedN];              // compacted (distinct, ascending) leaf indexes
    __shared__ double s_w[kSortedN];                  // their new leaf weights
    __shared__ short s_hb[kSortedN];
    __shared__ short s_next[kSortedN];
    __shared__ int s_warp[kSortedN / 32];
    __shared__ double s_red[kSortedN / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int levels = t.depth - 1;
    const bool in = tid < n;
    const int64_t k = in ? idx[tid] : 0;
    const bool oob = in && (k < 0 || k >= t.size);
    const bool unsorted = in && tid > 0 && idx[tid - 1] > k;
    // (__syncthreads_or returns a truth value, not the bitwise OR: two separate votes)
    const int any_oob = __syncthreads_or(oob ? 1 : 0);
    const int generic = __syncthreads_or((unsorted || valid != nullptr) ? 1 : 0);
    if (any_oob && !generic) {
        if (tid == 0) {
            t.st->err = MB_ERR_INDEX_OOB;
            t.st->err_gen = gen;
        }
        return;
    }
    if (generic) {   // also validates (a masked-out entry may legitimately be out of range)
        tree_update_body<WT, false>(t, idx, w, valid, n, nz, gen);
        return;
    }
    // leaf winners: the last entry of a run of equal keys (highest batch position); max over ALL entries
    const double x = in ? normalized_weight(w[tid], nz) : -INFINITY;
    const bool winner = in && (tid == n - 1 || idx[tid + 1] != k);
    double mx = mb::warp_max(x);
    const unsigned ballot = __ballot_sync(0xffffffffu, winner);
    if (lane == 0) {
        s_warp[wid] = __popc(ballot);
        s_red[wid] = mx;
    }
    __syncthreads();
    int base = 0, m = 0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) {
        if (j < wid) base += s_warp[j];
        m += s_warp[j];
    }
    if (tid == 0) {
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j) mx = fmax(mx, s_red[j]);
        if (mx > -INFINITY && !(t.st->max_leaf > mx)) {
            t.st->max_leaf = mx;
            t.st->max_is_f32 = sizeof(WT) == 4 ? 1 : 0;
        }
    }
    if (winner) {
        const int pos = base + __popc(ballot & ((1u << lane) - 1u));
        s_key[pos] = (unsigned)k;
        s_w[pos] = x;
    }
    __syncthreads();
    // from here on thread `tid` owns compacted entry `tid` (tid < m)
    bool alive = tid < m;
    unsigned key = 0;
    double val = 0.0;
    int hb = -1;
    if (alive) {
        key = s_key[tid];
        val = s_w[tid];
        hb = tid == 0 ? levels : 31 - __clz(s_key[tid - 1] ^ key);  // representative of its ancestor up to level hb
        s_hb[tid] = (short)hb;
        s_next[tid] = (short)(tid + 1);
        const int lim = hb < levels ? hb : levels - 1;               // siblings needed for steps 0..lim
        double pre[kSortedLevels];
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) pre[l] = __ldcg(t.W + level_off(t, l) + ((key >> l) ^ 1u));
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) park[l * kSortedN + tid] = pre[l];
    }
    const double leaf_val = val;
    unsigned wrote = 0;
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < levels; ++l) {
        if (alive && hb == l) {          // my node is a right child whose left sibling is touched: hand over, retire
            s_val[tid] = val;
            alive = false;
        }
        __syncthreads();
        if (alive) {
            const int j = s_next[tid];
            double other;
            if (j < m && s_hb[j] == l) {
                other = s_val[j];
                s_next[tid] = s_next[j];
            } else {
                other = park[l * kSortedN + tid];
            }
            val = val + other;           // left + right (IEEE addition is commutative)
            park[l * kSortedN + tid] = val;
            wrote |= 1u << l;
        }
    }
    if (tid < m) {
        t.W[key] = leaf_val;
        for (int l = 0; l < levels; ++l)
            if (wrote & (1u << l)) t.W[level_off(t, l + 1) + (key >> (l + 1))] = park[l * kSortedN + tid];
    }
}

// ---- second-generation small sorted update (opt-in until measured: MB_SORTED_CLUSTER=1, MB_SORTED_ASYNC=1) ----
// Two independent changes to tree_update_sorted, selected at compile time:
//
// CB = 3 (MB_SORTED_CLUSTER): the kernel above is bound by ONE SM's load/store unit: ~8.7 k scattered 8-byte
//   loads (stored siblings) and as many scattered stores for 512 entries of a 2^24-leaf tree, ~1 sector request
//   per cycle.  Here the tree is cut at level ls = levels - 3 into 8 subtrees; CTA r of a cluster of 8 owns the
//   entries whose leaf index lies in subtree r (a contiguous run of the sorted batch; entries of different
//   subtrees never meet below level ls), runs the walk for levels 0..ls-1 on its own SM -- prefetch and write-out
//   now use 8 load/store units -- and publishes the value of its subtree root (the stored one if it has no
//   entry) in shared memory; after cluster.sync() the leader reads the 8 roots over DSMEM and computes the 7
//   nodes above them.
//
// kAsync (MB_SORTED_ASYNC): the walk above spends one CTA-wide barrier per level although an entry only ever
//   waits for ONE other entry: who hands its value to whom is the Cartesian tree of hb[] (entry j hands the value
//   of its level-hb[j] node to the nearest entry on its left with a larger hb, which is the leftmost entry of the
//   sibling subtree).  So every entry walks its levels on its own -- add the prefetched stored sibling, or, when
//   its right-hand neighbour's hb equals the level, wait for that neighbour's published value and inherit its
//   neighbour pointer -- and publishes (value, next) with a flag when it reaches level hb.  No barrier in the walk;
//   the critical path is the ~log-depth chain of hand-overs.  Waiting is cooperative polling (every lane takes
//   one non-blocking step per round, a warp leaves the loop together), bounded: running out of rounds reports an
//   error instead of hanging.
//
// Parents stay fl(left + right) of the current children in both variants: bit-identical to the kernel above.
constexpr int kSpinLimit = 1 << 20;   // polling rounds per warp before the walk gives up (a few ms)

template <typename WT, int CB, bool kAsync>
__global__ void __launch_bounds__(kSortedN) tree_update_sorted_v2(TreeDev t, const int64_t *__restrict__ idx,
                                                                  const WT *__restrict__ w,
                                                                  const uint8_t *__restrict__ valid, int n,
                                                                  Normalize nz, int gen) {
    constexpr int kCtas = 1 << CB;
    MB_DYNAMIC_SMEM(dyn);
    double *park = reinterpret_cast<double *>(dyn);
    __shared__ double s_val[kSortedN];
    __shared__ unsigned s_key[kSortedN];
    __shared__ double s_w[kSortedN];
    __shared__ short s_hb[kSortedN];
    __shared__ short s_next[kSortedN];
    __shared__ int s_flag[kSortedN];
    __shared__ int s_warp[kSortedN / 32];
    __shared__ double s_red[kSortedN / 32];
    __shared__ double s_top;                          // value of this CTA's subtree root (level ls)
    __shared__ double s_max;                          // largest new leaf weight among this CTA's entries
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned rank = 0;
    if (CB > 0) rank = cg::this_cluster().block_rank();
    const int levels = t.depth - 1;
    const int ls = levels - CB;
    const bool in = tid < n;
    const int64_t k = in ? idx[tid] : 0;
    const bool oob = in && (k < 0 || k >= t.size);
    const bool unsorted = in && tid > 0 && idx[tid - 1] > k;
    if (kAsync) s_flag[tid] = 0;
    if (tid == 0) s_fail = 0;
    // every CTA looks at the whole batch, so all of them take the same branch (no cluster barrier is skipped)
    const int any_oob = __syncthreads_or(oob ? 1 : 0);
    const int generic = __syncthreads_or((unsorted || valid != nullptr) ? 1 : 0);
    if (any_oob && !generic) {
        if (rank == 0 && tid == 0) {
            t.st->err = MB_ERR_INDEX_OOB;
            t.st->err_gen = gen;
        }
        return;
    }
    if (generic) {   // CTA 0 alone: blockIdx.x == 0 and gridDim.x * blockDim.x >= n, so the body sees one CTA
        if (rank == 0) tree_update_body<WT, false>(t, idx, w, valid, n, nz, gen);
        return;
    }
    const bool mine = in && (unsigned)(k >> ls) == rank;
    const double x = mine ? normalized_weight(w[tid], nz) : -INFINITY;
    const bool winner = mine && (tid == n - 1 || idx[tid + 1] != k);
    double mx = mb::warp_max(x);
    const unsigned ballot = __ballot_sync(0xffffffffu, winner);
    if (lane == 0) {
        s_warp[wid] = __popc(ballot);
        s_red[wid] = mx;
    }
    __syncthreads();
    int base = 0, m = 0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) {
        if (j < wid) base += s_warp[j];
        m += s_warp[j];
    }
    if (tid == 0) {
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j) mx = fmax(mx, s_red[j]);
        s_max = mx;
    }
    if (winner) {
        const int pos = base + __popc(ballot & ((1u << lane) - 1u));
        s_key[pos] = (unsigned)k;
        s_w[pos] = x;
    }
    __syncthreads();
    bool alive = tid < m;
    unsigned key = 0;
    double val = 0.0;
    int hb = -1;
    if (alive) {
        key = s_key[tid];
        val = s_w[tid];
        hb = tid == 0 ? ls : 31 - __clz(s_key[tid - 1] ^ key);      // < ls for tid > 0: same subtree
        s_hb[tid] = (short)hb;
        s_next[tid] = (short)(tid + 1);
        const int lim = hb < ls ? hb : ls - 1;
        double pre[kSortedLevels];
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) pre[l] = __ldcg(t.W + level_off(t, l) + ((key >> l) ^ 1u));
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) park[l * kSortedN + tid] = pre[l];
    }
    const double leaf_val = val;
    unsigned wrote = 0;
    __syncthreads();
    if (kAsync) {
        // Cooperative polling, one warp at a time: in every round each lane that still has work takes ONE
        // non-blocking step (advance a level, or look at its neighbour's flag and try again next round), and the
        // warp leaves the loop together when all its lanes are done.  No lane ever spins while another lane of
        // its warp waits at a reconvergence point, so progress does not depend on how divergent lanes are scheduled.
        int nx = tid + 1, l = 0;
        const int lim = !alive ? 0 : (tid == 0 ? ls : hb);          // I compute my ancestors on levels 1..lim
        bool done = !alive;
        int rounds = 0;
        while (__any_sync(0xffffffffu, !done)) {
            if (!done) {
                if (l == lim) {                                  // reached the level where I hand over (entry 0: the top)
                    if (tid != 0) {
                        reinterpret_cast<volatile double *>(s_val)[tid] = val;
                        reinterpret_cast<volatile short *>(s_next)[tid] = (short)nx;
                        __threadfence_block();
                        reinterpret_cast<volatile int *>(s_flag)[tid] = 1;
                    }
                    done = true;
                } else if (nx < m && s_hb[nx] == l) {            // my right-hand neighbour ends here: take its node value
                    if (reinterpret_cast<volatile int *>(s_flag)[nx] != 0) {
                        __threadfence_block();
                        const double other = reinterpret_cast<volatile double *>(s_val)[nx];
                        nx = reinterpret_cast<volatile short *>(s_next)[nx];
                        val = val + other;                       // left + right
                        park[l * kSortedN + tid] = val;
                        wrote |= 1u << l;
                        ++l;
                    }                                            // else: not published yet, look again next round
                } else {
                    val = val + park[l * kSortedN + tid];        // stored sibling (prefetched)
                    park[l * kSortedN + tid] = val;
                    wrote |= 1u << l;
                    ++l;
                }
            }
            if (++rounds > kSpinLimit) {                         // uniform per warp; cannot happen unless the logic is broken
                if (!done) s_fail = 1;
                break;
            }
        }
        __syncthreads();
        if (s_fail) {                                            // uniform: report, write nothing
            if (tid == 0) {
                t.st->err = MB_ERR_INVALID;
                t.st->err_gen = gen;
            }
            wrote = 0;
            alive = false;
        }
    } else {
#pragma unroll 1
        for (int l = 0; l < ls; ++l) {
            if (alive && hb == l) {
                s_val[tid] = val;
                alive = false;
            }
            __syncthreads();
            if (alive) {
                const int j = s_next[tid];
                double other;
                if (j < m && s_hb[j] == l) {
                    other = s_val[j];
                    s_next[tid] = s_next[j];
                } else {
                    other = park[l * kSortedN + tid];
                }
                val = val + other;
                park[l * kSortedN + tid] = val;
                wrote |= 1u << l;
            }
        }
    }
    const bool failed = kAsync && s_fail;
    if (tid == 0) s_top = (m > 0 && !failed) ? val : __ldcg(t.W + level_off(t, ls) + rank);
    if (tid < m && !failed) {
        t.W[key] = leaf_val;
        for (int l = 0; l < ls; ++l)
            if (wrote & (1u << l)) t.W[level_off(t, l + 1) + (key >> (l + 1))] = park[l * kSortedN + tid];
    }
    if (CB > 0) cg::this_cluster().sync();
    else __syncthreads();
    if (rank == 0 && tid == 0) {
        double v[kCtas], bm = -INFINITY;
#pragma unroll
        for (int r = 0; r < kCtas; ++r) {
            if (CB > 0) {
                v[r] = *cg::this_cluster().map_shared_rank(&s_top, r);
                bm = fmax(bm, *cg::this_cluster().map_shared_rank(&s_max, r));
            } else {
                v[r] = s_top;
                bm = s_max;
            }
        }
        int cnt = kCtas;
#pragma unroll
        for (int b = 1; b <= CB; ++b) {
            cnt >>= 1;
#pragma unroll
            for (int j = 0; j < (kCtas + 1) / 2; ++j)
                if (j < cnt) {
                    v[j] = v[2 * j] + v[2 * j + 1];
                    t.W[level_off(t, ls + b) + j] = v[j];
                }
        }
        if (bm > -INFINITY && !(t.st->max_leaf > bm)) {
            t.st->max_leaf = bm;
            t.st->max_is_f32 = sizeof(WT) == 4 ? 1 : 0;
        }
    }
    if (CB > 0) cg::this_cluster().sync();   // nobody exits (and frees its shared memory) while the leader may still read it
}


// ---- third generation: the update spread over the whole GPU in ONE launch, no grid barrier ---------------------
// The single-CTA kernels above are bound by one SM's load/store unit (~8.7 k scattered sibling loads and as many
// scattered stores for 512 entries of a 2^24-leaf tree) and by 24 CTA-wide barriers.  Here the tree is cut at level
// ls = levels - cut into G = 2^cut subtrees, ONE CTA PER SUBTREE (G = 128 at 10 M leaves: the prefetch and the
// write-out run on 128 load/store units and every CTA's sibling loads are in flight at the same time):
//   A  every CTA scans the whole batch (n <= 4096 indexes: 32 KB out of L2) and keeps the entries of its own
//      subtree -- ordered compaction, so a sorted batch stays sorted; validation sees the whole batch in every CTA,
//      so all CTAs take the same decision without talking to each other
//   B  entries of different subtrees never meet below level ls: the CTA sorts its entries if they are not
//      ascending (bitonic, shared memory), keeps the LAST of equal keys (highest batch position: last write wins),
//      and runs the static merge-tree walk of tree_update_sorted for levels 0..ls-1 (a warp-only walk with
//      __syncwarp when it has <= 32 entries, which is the normal case: 512 draws / 128 subtrees)
//   C  __threadfence + a ticket: the CTA that draws the last ticket re-derives the G - 1 nodes above level ls from
//      the G subtree roots (untouched roots are re-added to the value they already hold, by the tree invariant),
//      folds the batch maximum into max_leaf and resets the ticket for the next launch.
// A CTA with more than kPartCap distinct entries (only possible for n > 512 on a heavily skewed batch) walks its
// levels through global memory like the generic body -- slower, still correct, still independent of the others.
// Parents are fl(left + right) of the current children everywhere: bit-identical to the reference update.
#ifndef MB_PART_THREADS
#define MB_PART_THREADS 512         // (the CPU emulation harness compiles the same source with 128: 4x fewer OS threads)
#endif
constexpr int kPartThreads = MB_PART_THREADS;
constexpr int kPartCap = kPartThreads;   // distinct entries a CTA walks out of shared memory (one thread per entry)
constexpr int kPartMaxN = 4096;     // largest batch served by this kernel
constexpr int kPartMaxCut = 7;      // at most 128 subtrees
constexpr int kPartPosBits = 12;    // batch position inside the 64-bit sort key (kPartMaxN = 2^12)

__host__ __device__ __forceinline__ int part_cut(int levels) {
    const int c = levels - 6;
    return c < 0 ? 0 : (c > kPartMaxCut ? kPartMaxCut : c);
}
__host__ __device__ __forceinline__ size_t part_dyn_smem(int levels) {
    return sizeof(unsigned long long) * kPartMaxN + sizeof(double) * (size_t)kPartCap * (size_t)(levels - part_cut(levels) + 1);
}

// The merge-tree walk of one subtree.  kWarp: the caller is warp 0 alone and the subtree has <= 32 entries.
template <bool kWarp>
__device__ __forceinline__ void part_walk(const TreeDev &t, double *park, double *s_val, const unsigned *s_key,
                                          const double *s_w, short *s_hb, short *s_next, int m, int ls, int tid) {
    bool alive = tid < m;
    unsigned key = 0;
    double val = 0.0;
    int hb = -1;
    if (alive) {
        key = s_key[tid];
        val = s_w[tid];
        hb = tid == 0 ? ls : 31 - __clz(s_key[tid - 1] ^ key);      // < ls for tid > 0: same subtree
        s_hb[tid] = (short)hb;
        s_next[tid] = (short)(tid + 1);
        const int lim = hb < ls ? hb : ls - 1;
        double pre[kSortedLevels];
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) pre[l] = __ldcg(t.W + level_off(t, l) + ((key >> l) ^ 1u));
#pragma unroll
        for (int l = 0; l < kSortedLevels; ++l)
            if (l <= lim) park[l * kPartCap + tid] = pre[l];
    }
    const double leaf_val = val;
    unsigned wrote = 0;
    if (kWarp) __syncwarp();
    else __syncthreads();
#pragma unroll 1
    for (int l = 0; l < ls; ++l) {
        if (alive && hb == l) {          // right child whose left sibling is touched too: hand over, retire
            s_val[tid] = val;
            alive = false;
        }
        if (kWarp) __syncwarp();
        else __syncthreads();
        if (alive) {
            const int j = s_next[tid];
            double other;
            if (j < m && s_hb[j] == l) {
                other = s_val[j];
                s_next[tid] = s_next[j];
            } else {
                other = park[l * kPartCap + tid];
            }
            val = val + other;           // left + right (IEEE addition is commutative)
            park[l * kPartCap + tid] = val;
            wrote |= 1u << l;
        }
    }
    if (tid < m) {
        t.W[key] = leaf_val;
        for (int l = 0; l < ls; ++l)
            if (wrote & (1u << l)) t.W[level_off(t, l + 1) + (key >> (l + 1))] = park[l * kPartCap + tid];
    }
}

template <typename WT>
__global__ void __launch_bounds__(kPartThreads) tree_update_part(TreeDev t, const int64_t *__restrict__ idx,
                                                                 const WT *__restrict__ w,
                                                                 const uint8_t *__restrict__ valid, int n,
                                                                 Normalize nz, int gen) {
    MB_DYNAMIC_SMEM(dyn);
    unsigned long long *s_ent = reinterpret_cast<unsigned long long *>(dyn);      // [kPartMaxN] (leaf << 12) | position
    double *park = reinterpret_cast<double *>(dyn + sizeof(unsigned long long) * kPartMaxN);   // [ls][kPartCap]
    __shared__ double s_val[kPartCap];
    __shared__ unsigned s_key[kPartCap];
    __shared__ double s_w[kPartCap];
    __shared__ short s_hb[kPartCap];
    __shared__ short s_next[kPartCap];
    __shared__ int s_warp[kPartThreads / 32];
    __shared__ double s_red[kPartThreads / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kPartThreads / 32;
    const int levels = t.depth - 1;
    const int cut = part_cut(levels);
    const int ls = levels - cut;
    const unsigned rank = blockIdx.x;
    // ---- A: scan the whole batch, keep my subtree's entries in batch order
    int m = 0;
    bool oob = false;
    double mx = -INFINITY;
    for (int base = 0; base < n; base += kPartThreads) {
        const int i = base + tid;
        bool mine = false;
        int64_t k = 0;
        if (i < n && (!valid || valid[i])) {
            k = idx[i];
            if (k < 0 || k >= t.size) oob = true;
            else mine = (unsigned)(k >> ls) == rank;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) s_warp[wid] = __popc(ballot);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int j = 0; j < kWarps; ++j) {
            if (j < wid) before += s_warp[j];
            tot += s_warp[j];
        }
        if (mine) {
            s_ent[m + before + __popc(ballot & ((1u << lane) - 1u))] =
                ((unsigned long long)k << kPartPosBits) | (unsigned long long)i;
            mx = fmax(mx, normalized_weight(w[i], nz));      // max over ALL written weights, duplicates included (:175)
        }
        m += tot;
        __syncthreads();
    }
    if (__syncthreads_or(oob ? 1 : 0)) {   // every CTA sees the same batch: all of them leave, nothing is written
        if (rank == 0 && tid == 0) {
            t.st->err = MB_ERR_INDEX_OOB;
            t.st->err_gen = gen;
        }
        return;
    }
    if (m > 0) {   // (uniform per CTA)
        mx = mb::warp_max(mx);
        if (lane == 0) s_red[wid] = mx;
        // ---- B1: ascending?  (a batch that came out of the stratified draw is; merged shard blocks are not)
        bool unsorted = false;
        for (int i = tid + 1; i < m; i += kPartThreads) unsorted |= s_ent[i - 1] > s_ent[i];
        if (__syncthreads_or(unsorted ? 1 : 0)) {
            int p2 = 1;
            while (p2 < m) p2 <<= 1;
            for (int i = m + tid; i < p2; i += kPartThreads) s_ent[i] = ~0ull;
            __syncthreads();
            for (int k2 = 2; k2 <= p2; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < p2; i += kPartThreads) {
                        const int q = i ^ j;
                        if (q > i) {
                            const unsigned long long a = s_ent[i], b = s_ent[q];
                            if ((a > b) == ((i & k2) == 0)) {
                                s_ent[i] = b;
                                s_ent[q] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
        }
        if (tid == 0) {
            double bm = s_red[0];
            for (int j = 1; j < kWarps; ++j) bm = fmax(bm, s_red[j]);
            if (bm > -INFINITY) atomic_max_f64(&t.st->batch_max, bm);
        }
        // ---- B2: distinct leaves (the last of equal keys carries the highest batch position)
        int distinct = 0;
        for (int base = 0; base < m; base += kPartThreads) {
            const int i = base + tid;
            const bool win = i < m && (i == m - 1 || (s_ent[i + 1] >> kPartPosBits) != (s_ent[i] >> kPartPosBits));
            const unsigned ballot = __ballot_sync(0xffffffffu, win);
            if (lane == 0) s_warp[wid] = __popc(ballot);
            __syncthreads();
            int before = 0, tot = 0;
#pragma unroll
            for (int j = 0; j < kWarps; ++j) {
                if (j < wid) before += s_warp[j];
                tot += s_warp[j];
            }
            const int pos = distinct + before + __popc(ballot & ((1u << lane) - 1u));
            if (win && pos < kPartCap) {
                const unsigned long long e = s_ent[i];
                s_key[pos] = (unsigned)(e >> kPartPosBits);
                s_w[pos] = normalized_weight(w[(int)(e & ((1u << kPartPosBits) - 1u))], nz);
            }
            distinct += tot;
            __syncthreads();
        }
        if (distinct <= 32) {
            if (wid == 0) part_walk<true>(t, park, s_val, s_key, s_w, s_hb, s_next, distinct, ls, tid);
        } else if (distinct <= kPartCap) {
            part_walk<false>(t, park, s_val, s_key, s_w, s_hb, s_next, distinct, ls, tid);
        } else {
            // skewed large batch: level-synchronous recompute through global memory, my entries only
            for (int i = tid; i < m; i += kPartThreads) {
                const unsigned long long e = s_ent[i];
                if (i == m - 1 || (s_ent[i + 1] >> kPartPosBits) != (e >> kPartPosBits))
                    t.W[e >> kPartPosBits] = normalized_weight(w[(int)(e & ((1u << kPartPosBits) - 1u))], nz);
            }
            for (int l = 1; l <= ls; ++l) {
                __syncthreads();
                const double *below = t.W + level_off(t, l - 1);
                double *here = t.W + level_off(t, l);
                for (int i = tid; i < m; i += kPartThreads) {
                    const int64_t node = (int64_t)(s_ent[i] >> kPartPosBits) >> l;
                    const double2 c = __ldcg(reinterpret_cast<const double2 *>(below + 2 * node));
                    here[node] = c.x + c.y;
                }
            }
        }
    }
    // ---- C: ticket; the last CTA lifts the levels above the cut
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&t.st->ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int G = 1 << cut;
    if (tid < G) s_val[tid] = __ldcg(t.W + level_off(t, ls) + tid);
    for (int b = 1; b <= cut; ++b) {
        __syncthreads();
        double v = 0.0;
        const bool on = tid < (G >> b);
        if (on) v = s_val[2 * tid] + s_val[2 * tid + 1];
        __syncthreads();
        if (on) {
            s_val[tid] = v;
            t.W[level_off(t, ls + b) + tid] = v;
        }
    }
    if (tid == 0) {
        // max_leaf = max(np.max(weights), max_leaf): the new value (and its dtype) wins ties (:175)
        const double bm = __ldcg(&t.st->batch_max);
        if (bm > -INFINITY && !(t.st->max_leaf > bm)) {
            t.st->max_leaf = bm;
            t.st->max_is_f32 = sizeof(WT) == 4 ? 1 : 0;
        }
        t.st->batch_max = -INFINITY;
        t.st->ticket = 0;
    }
}
