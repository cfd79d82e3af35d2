// Shard exchange over peer memory (NVLink 5 / NVSwitch) for sm_100a: entry points + the cross-GPU barrier.
//
// The fixed-shape APEX update moves three things between the GPUs of a box every update: the sampled rows
// (prioritized_buffer_d.py:236-266), the priority write-backs (:184-202) and the gradients.  Over NCCL each is a
// collective launch with its own internal synchronisation (~15-25 us for these message sizes); here the producing
// kernel stores straight into the consumer's buffer (every GPU's buffer is mapped into every process once, at set-up)
// and ONE small barrier kernel per exchange orders producer and consumer:
//   every member: fence(system); flag[self] of every member <- epoch (release)   -- "my pushes are done"
//                 wait until all of my own flags reached epoch (acquire)          -- "everybody's pushes to me are done"
// The epoch lives in device memory and advances inside the kernel, so a captured graph replays the barrier unchanged.
// A member that never arrives (crashed process) would leave the others spinning: the wait gives up after timeout_ns
// and raises d_state[1], which the host side checks -- a stuck exchange must not hold a GPU box.
#include <cuda/atomic>

#include "mb_common.cuh"

namespace {

#include "rows_device.cuh"   // PeerPtrs, wire_push_rows_kernel, wire_push_update_kernel, peer_push_kernel, peer_sum_kernel

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerPtrs flags, int P, int self_block,
                                                          unsigned long long *state, long long timeout_ns) {
    const int r = threadIdx.x;
    const unsigned long long epoch = state[0] + 1;       // (every lane reads the same value; lane 0 writes it back last)
    __syncwarp();
    if (r < P) {
        // release at system scope: everything this GPU wrote before the barrier kernel started (stream order) and
        // inside it is visible to whoever acquires the flag
        cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> out(
            *(reinterpret_cast<unsigned long long *>(flags.p[r]) + self_block));
        out.store(epoch, cuda::memory_order_release);
        cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> in(
            *(reinterpret_cast<unsigned long long *>(flags.p[self_block]) + r));
        const unsigned long long t0 = global_ns();
        while (in.load(cuda::memory_order_acquire) < epoch) {
            if ((long long)(global_ns() - t0) > timeout_ns) {
                state[1] = 1;                              // a member did not arrive: reported by the host side
                break;
            }
            __nanosleep(64);
        }
    }
    __syncwarp();
    if (r == 0) state[0] = epoch;
}

// shard sums: all-gather + barrier + apportioning in ONE warp.  Lane r stores this member's sum into slot `self` of
// member r (payload first, then the flag with release semantics), waits for member r's flag, then lane 0 apportions.
__global__ void __launch_bounds__(32) peer_sums_apportion_kernel(PeerPtrs flags, PeerPtrs sums, const double *my_sum, int P,
                                                                 int self_block, unsigned long long *state, long long timeout_ns,
                                                                 int64_t batch, int64_t merged_rows, const int64_t *n_stored,
                                                                 double *beta, double beta_increment, int64_t *counts,
                                                                 double *scale, mb_per_params *params) {
    const int r = threadIdx.x;
    const unsigned long long epoch = state[0] + 1;
    __syncwarp();
    if (r < P) {
        reinterpret_cast<volatile double *>(sums.p[r])[self_block] = *my_sum;
        cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> out(
            *(reinterpret_cast<unsigned long long *>(flags.p[r]) + self_block));
        out.store(epoch, cuda::memory_order_release);
        cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> in(
            *(reinterpret_cast<unsigned long long *>(flags.p[self_block]) + r));
        const unsigned long long t0 = global_ns();
        while (in.load(cuda::memory_order_acquire) < epoch) {
            if ((long long)(global_ns() - t0) > timeout_ns) {
                state[1] = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncwarp();
    if (r == 0) {
        state[0] = epoch;
        wire_apportion_body(reinterpret_cast<const double *>(sums.p[self_block]), P, batch, self_block, merged_rows, n_stored, beta,
                            beta_increment, counts, scale, params);
    }
}

int fill(PeerPtrs &pp, void *const *peers, int n_blocks) {
    if (!peers || n_blocks < 1 || n_blocks > MB_MAX_PEERS) return MB_ERR_INVALID;
    for (int m = 0; m < MB_MAX_PEERS; ++m) pp.p[m] = m < n_blocks ? (char *)peers[m] : nullptr;
    for (int m = 0; m < n_blocks; ++m)
        if (!pp.p[m]) return MB_ERR_INVALID;
    return MB_OK;
}

// signal / wait descriptor of one site; flags == NULL: off
int fill_signal(PeerSignal &sg, void *const *peer_flags, int n_blocks, int self_block, uint64_t *d_state, unsigned int *d_ticket) {
    sg = PeerSignal{};
    if (!peer_flags) return MB_OK;
    const int r = fill(sg.flags, peer_flags, n_blocks);
    if (r != MB_OK) return r;
    if (!d_state || n_blocks > 32) return MB_ERR_INVALID;
    sg.state = (unsigned long long *)d_state;
    sg.ticket = d_ticket;
    sg.P = n_blocks;
    sg.self_block = self_block;
    sg.on = 1;
    return MB_OK;
}

}  // namespace

extern "C" {

int mb_wire_push_rows(const void *d_ring, int64_t pitch, int64_t capacity, const int64_t *d_index,
                      const double *d_is_weight, const int64_t *d_version_table, int n_blocks, int64_t cap,
                      int self_block, int64_t width, void *const *peers, void *const *peer_flags, unsigned int *d_ticket,
                      uint64_t *d_state, void *stream) {
    PeerPtrs pp;
    int r = fill(pp, peers, n_blocks);
    if (r != MB_OK) return r;
    PeerSignal sg;
    r = fill_signal(sg, peer_flags, n_blocks, self_block, d_state, d_ticket);
    if (r != MB_OK || (peer_flags && !d_ticket)) return r != MB_OK ? r : MB_ERR_INVALID;
    if (!d_ring || !d_index || !d_is_weight || !d_version_table || pitch <= 0 || capacity <= 0 || cap < 1 ||
        self_block < 0 || self_block >= n_blocks || width % 8 || pitch % 8 || width < pitch + 24)
        return MB_ERR_INVALID;
    const int64_t rows = (int64_t)n_blocks * cap;
    int64_t blocks = (rows + 7) / 8;                       // 8 warps (rows) per CTA
    const int64_t most = (int64_t)mb::kNumSMs * 16;
    if (blocks > most) blocks = most;
    wire_push_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_ring, pitch, capacity, d_index, d_is_weight, d_version_table, n_blocks, cap, self_block, width, pp, sg);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_push_update(int n_blocks, int64_t cap, int self_block, const int64_t *d_counts, const int64_t *d_index,
                        const int64_t *d_version, const float *d_priority, int64_t merged_rows, void *const *peers,
                        void *stream) {
    PeerPtrs pp;
    const int r = fill(pp, peers, n_blocks);
    if (r != MB_OK) return r;
    if (!d_counts || !d_index || !d_version || !d_priority || cap < 1 || self_block < 0 || self_block >= n_blocks ||
        merged_rows < 0)
        return MB_ERR_INVALID;
    if (merged_rows == 0) return MB_OK;
    wire_push_update_kernel<<<(unsigned)((merged_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_blocks, cap, self_block, d_counts, d_index, d_version, d_priority, merged_rows, pp);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_peer_push(const void *d_src, int64_t nbytes, int n_blocks, int self_block, int64_t slot_stride,
                 void *const *peers, void *const *peer_flags, unsigned int *d_ticket, uint64_t *d_state, void *stream) {
    PeerPtrs pp;
    int r = fill(pp, peers, n_blocks);
    if (r != MB_OK) return r;
    PeerSignal sg;
    r = fill_signal(sg, peer_flags, n_blocks, self_block, d_state, d_ticket);
    if (r != MB_OK || (peer_flags && !d_ticket)) return r != MB_OK ? r : MB_ERR_INVALID;
    if (!d_src || nbytes < 0 || slot_stride < nbytes || self_block < 0 || self_block >= n_blocks) return MB_ERR_INVALID;
    if (nbytes == 0) return MB_OK;
    int64_t bx = (nbytes / 16 + 255) / 256;
    bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
    peer_push_kernel<<<dim3((unsigned)bx, (unsigned)n_blocks), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_src, nbytes, self_block, slot_stride, pp, sg);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_peer_sum(const float *d_slots, int n_blocks, int64_t n, int64_t slot_floats, float divisor, float *d_out,
                void *const *peer_flags, int self_block, uint64_t *d_state, void *stream) {
    if (!d_slots || !d_out || n_blocks < 1 || n < 0 || slot_floats < n || divisor == 0.0f) return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    PeerSignal wait;
    const int r = fill_signal(wait, peer_flags, n_blocks, self_block, d_state, nullptr);
    if (r != MB_OK) return r;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 2 * mb::kNumSMs) blocks = 2 * mb::kNumSMs;
    peer_sum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_slots, n_blocks, n, slot_floats, divisor, d_out, wait);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_peer_barrier(void *const *peer_flags, int n_blocks, int self_block, uint64_t *d_state,
                    int64_t timeout_ns, void *stream) {
    PeerPtrs pp;
    const int r = fill(pp, peer_flags, n_blocks);
    if (r != MB_OK) return r;
    if (!d_state || self_block < 0 || self_block >= n_blocks || n_blocks > 32 || timeout_ns <= 0) return MB_ERR_INVALID;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, n_blocks, self_block, (unsigned long long *)d_state, (long long)timeout_ns);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_compact_wait(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                         const int64_t *d_counts, void *d_merged, int64_t merged_rows, int64_t *d_index, int64_t *d_version,
                         double *d_is_weight, const double *d_scale, void *const *peer_flags, int self_block,
                         uint64_t *d_state, void *stream) {
    if (!d_recv || !d_counts || !d_merged || !d_index || !d_version || !d_is_weight || n_blocks < 1 || cap < 1 ||
        width % 8 || meta_offset % 8 || meta_offset + 24 > width || merged_rows < 0)
        return MB_ERR_INVALID;
    if (merged_rows == 0) return MB_OK;
    PeerSignal wait;
    const int r = fill_signal(wait, peer_flags, n_blocks, self_block, d_state, nullptr);
    if (r != MB_OK) return r;
    wire_compact_kernel<<<(unsigned)((merged_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_recv, n_blocks, cap, width, meta_offset, d_counts, (char *)d_merged, merged_rows, d_index,
        d_version, d_is_weight, d_scale, wait);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_compact_fields(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                           const int64_t *d_counts, const mb_field *fields, int n_fields, int64_t merged_rows,
                           int64_t *d_index, int64_t *d_version, double *d_is_weight, const double *d_scale,
                           void *const *peer_flags, int self_block, uint64_t *d_state, void *stream) {
    if (!d_recv || !d_counts || !fields || n_fields < 1 || n_fields > MB_MAX_FIELDS || !d_index || !d_version || !d_is_weight ||
        n_blocks < 1 || cap < 1 || width % 8 || meta_offset % 8 || meta_offset + 24 > width || merged_rows < 0)
        return MB_ERR_INVALID;
    if (merged_rows == 0) return MB_OK;
    FieldList fl;
    fl.n = n_fields;
    for (int i = 0; i < n_fields; ++i) {
        fl.f[i] = fields[i];
        if (fields[i].row_offset < 0 || fields[i].nbytes < 0 || fields[i].row_offset + fields[i].nbytes > meta_offset || !fields[i].d_dense)
            return MB_ERR_INVALID;
    }
    PeerSignal wait;
    const int r = fill_signal(wait, peer_flags, n_blocks, self_block, d_state, nullptr);
    if (r != MB_OK) return r;
    wire_compact_fields_kernel<<<(unsigned)((merged_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_recv, n_blocks, cap, width, meta_offset, d_counts, fl, merged_rows, d_index, d_version, d_is_weight,
        d_scale, wait);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_peer_sums_apportion(void *const *peer_flags, void *const *peer_sums, const double *d_my_sum, int n_blocks,
                           int self_block, uint64_t *d_state, int64_t timeout_ns, int64_t batch, int64_t merged_rows,
                           const int64_t *d_n_stored, double *d_beta, double beta_increment, int64_t *d_counts,
                           double *d_scale, mb_per_params *d_params, void *stream) {
    PeerPtrs fl, sm;
    int r = fill(fl, peer_flags, n_blocks);
    if (r != MB_OK) return r;
    r = fill(sm, peer_sums, n_blocks);
    if (r != MB_OK) return r;
    if (!d_my_sum || !d_state || self_block < 0 || self_block >= n_blocks || n_blocks > 32 || timeout_ns <= 0 || batch < 1 ||
        merged_rows < 1 || !d_n_stored || !d_beta || !d_counts || !d_scale || !d_params)
        return MB_ERR_INVALID;
    peer_sums_apportion_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fl, sm, d_my_sum, n_blocks, self_block,
                                                                   (unsigned long long *)d_state, (long long)timeout_ns, batch,
                                                                   merged_rows, d_n_stored, d_beta, beta_increment, d_counts,
                                                                   d_scale, d_params);
    MB_LAUNCHED(1);
    return MB_OK;
}

}  // extern "C"
