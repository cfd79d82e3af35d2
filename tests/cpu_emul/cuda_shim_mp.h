// TEST INFRASTRUCTURE ONLY -- a larger shim of the CUDA execution model for kernels that use warp intrinsics,
// atomics and thread-block clusters: every CTA of the (single) cluster is a forked PROCESS, every CUDA thread an
// OS thread of it; "global memory" is a MAP_SHARED region set up before the fork; function-scope __shared__
// variables are plain statics of the CTA's process; cluster.sync() is a cross-process barrier and
// cluster.map_shared_rank() fetches the bytes from the owning process, where a server thread answers such requests.  Slow and simple: it exists to run a kernel's DEVICE SOURCE in the build container (no GPU) and
// compare against the oracle.  Nothing under machin_b200/ includes it.
#pragma once
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static                       /* one CTA per process */
#define MB_DYNAMIC_SMEM(name) unsigned char *name = shim::dyn_smem

struct shim_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local shim_dim3 threadIdx;
inline shim_dim3 blockIdx, blockDim, gridDim;   // per process

struct alignas(16) double2 { double x, y; };

namespace shim {

constexpr int kMaxCtas = 8;
struct Mailbox {                                // one per CTA, in shared memory
    std::atomic<int> lock;                      // one requester at a time (any thread of any other CTA)
    std::atomic<int> request;                   // 0 idle, 1 pending, 2 answered
    const void *addr;
    int bytes;
    unsigned char reply[64];
};
struct ClusterShared {
    std::atomic<int> arrived[4];                // one counter per cluster.sync() generation (mod 4)
    Mailbox box[kMaxCtas];
};

inline ClusterShared *cluster = nullptr;
inline int n_ctas = 1, rank = 0, sync_gen = 0;
inline unsigned char *dyn_smem = nullptr;
inline std::barrier<> *cta_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
struct WarpScratch { std::atomic<unsigned> bits{0}; double slot[32]; };
inline std::vector<WarpScratch> warp_scratch;
inline std::atomic<int> cta_or{0};

inline void *shared_alloc(size_t bytes) {
    void *p = mmap(nullptr, bytes ? bytes : 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { std::perror("mmap"); std::exit(3); }
    return p;
}

inline void serve_mailbox() {
    Mailbox &m = cluster->box[rank];
    if (m.request.load(std::memory_order_acquire) == 1) {
        std::memcpy(m.reply, m.addr, (size_t)m.bytes);
        m.request.store(2, std::memory_order_release);
    }
}

}  // namespace shim

// ---------------------------------------------------------------- CTA-level
inline void __syncthreads() { shim::cta_barrier->arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
    if (pred) shim::cta_or.store(1);
    __syncthreads();
    const int r = shim::cta_or.load();
    __syncthreads();
    if (threadIdx.x == 0) shim::cta_or.store(0);
    __syncthreads();
    return r;
}
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---------------------------------------------------------------- warp-level (all 32 lanes take part)
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    shim::WarpScratch &s = shim::warp_scratch[w];
    if (pred) s.bits.fetch_or(1u << lane);
    shim::warp_barrier[w]->arrive_and_wait();
    const unsigned r = s.bits.load();
    shim::warp_barrier[w]->arrive_and_wait();
    if (lane == 0) s.bits.store(0);
    shim::warp_barrier[w]->arrive_and_wait();
    return r;
}
inline void __syncwarp() { shim::warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }   // all 32 lanes of the warp call it
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline double __shfl_xor_sync(unsigned, double v, int o) {
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    shim::WarpScratch &s = shim::warp_scratch[w];
    s.slot[lane] = v;
    shim::warp_barrier[w]->arrive_and_wait();
    const double r = s.slot[lane ^ (unsigned)o];
    shim::warp_barrier[w]->arrive_and_wait();
    return r;
}
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
using std::max;
using std::min;
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }

// ---------------------------------------------------------------- memory
template <typename T> inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <> inline double2 __ldcg(const double2 *p) { double2 v; v.x = p->x; v.y = p->y; return v; }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
inline int atomicMax(int *a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicAdd(int *a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicCAS(unsigned long long *a, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

// ---------------------------------------------------------------- cooperative groups
namespace cooperative_groups {
struct grid_group { void sync() const { std::abort(); } };           // cooperative grids are not emulated
inline grid_group this_grid() { return {}; }
struct cluster_group {
    unsigned block_rank() const { return (unsigned)shim::rank; }
    unsigned num_blocks() const { return (unsigned)shim::n_ctas; }
    void sync() const {
        __syncthreads();
        if (threadIdx.x == 0) {
            std::atomic<int> &c = shim::cluster->arrived[shim::sync_gen & 3];
            c.fetch_add(1);
            while (c.load(std::memory_order_acquire) < shim::n_ctas) std::this_thread::yield();
            shim::sync_gen++;
        }
        __syncthreads();
    }
    template <typename T> T *map_shared_rank(T *p, unsigned r) const {
        static_assert(sizeof(T) <= 64, "mailbox reply too small");
        if ((int)r == shim::rank) return p;
        static thread_local T copy;
        shim::Mailbox &m = shim::cluster->box[r];
        int expected = 0;
        while (!m.lock.compare_exchange_weak(expected, 1, std::memory_order_acquire)) {
            expected = 0;
            std::this_thread::yield();
        }
        m.addr = p;                                  // same binary, same address in every CTA's process
        m.bytes = (int)sizeof(T);
        m.request.store(1, std::memory_order_release);
        while (m.request.load(std::memory_order_acquire) != 2) std::this_thread::yield();
        std::memcpy(&copy, m.reply, sizeof(T));
        m.request.store(0, std::memory_order_release);
        m.lock.store(0, std::memory_order_release);
        return &copy;
    }
};
inline cluster_group this_cluster() { return {}; }
}  // namespace cooperative_groups
namespace cg = cooperative_groups;

namespace mb {
inline double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
}  // namespace mb

namespace shim {

// Launch one cluster of `ctas` CTAs of `block` threads (block % 32 == 0).  Returns 0 if every CTA exited cleanly.
inline int launch(int ctas, unsigned block, size_t dyn_bytes, const std::function<void()> &kernel, int clusters = 1);

inline int launch_one_cluster(int ctas, unsigned block, size_t dyn_bytes, const std::function<void()> &kernel,
                              int first_block, int grid_blocks) {
    cluster = static_cast<ClusterShared *>(shared_alloc(sizeof(ClusterShared)));
    new (cluster) ClusterShared();
    n_ctas = ctas;
    std::vector<pid_t> kids;
    for (int c = 0; c < ctas; ++c) {
        const pid_t pid = fork();
        if (pid < 0) { std::perror("fork"); return 4; }
        if (pid == 0) {
            rank = c;
            blockIdx.x = (unsigned)(first_block + c);
            blockDim.x = block;
            gridDim.x = (unsigned)grid_blocks;
            dyn_smem = static_cast<unsigned char *>(std::aligned_alloc(16, dyn_bytes ? (dyn_bytes + 15) / 16 * 16 : 16));
            std::barrier<> bar((std::ptrdiff_t)block);
            cta_barrier = &bar;
            const unsigned warps = block / 32;
            warp_scratch = std::vector<WarpScratch>(warps);
            for (unsigned w = 0; w < warps; ++w) warp_barrier.emplace_back(new std::barrier<>(32));
            // distributed shared memory: a server thread answers reads of this CTA's statics for as long as the CTA
            // lives (CUDA kernels end with a cluster.sync() so that no CTA exits while a peer may still read it)
            std::atomic<bool> stop{false};
            std::thread server([&] {
                while (!stop.load(std::memory_order_acquire)) {
                    serve_mailbox();
                    std::this_thread::yield();
                }
            });
            std::vector<std::thread> threads;
            threads.reserve(block);
            for (unsigned t = 0; t < block; ++t)
                threads.emplace_back([&, t] { threadIdx.x = t; kernel(); });
            for (auto &th : threads) th.join();
            stop.store(true, std::memory_order_release);
            server.join();
            std::_Exit(0);
        }
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 5;
    }
    return bad;
}

// `clusters` independent clusters of `ctas` CTAs each (grid = clusters * ctas blocks), one cluster after the other.
inline int launch(int ctas, unsigned block, size_t dyn_bytes, const std::function<void()> &kernel, int clusters) {
    for (int cl = 0; cl < clusters; ++cl) {
        const int rc = launch_one_cluster(ctas, block, dyn_bytes, kernel, cl * ctas, clusters * ctas);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace shim
