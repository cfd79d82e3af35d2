// Replay-row movement and shard-exchange kernels: device code only (no CUDA runtime calls).  Included by rows.cu
// inside its anonymous namespace and by tests/cpu_emul/wire_harness.cpp, which compiles THIS SOURCE with g++
// against a shim of the CUDA execution model so that the fixed-shape exchange can be replayed for 4 and 8 shards
// in the build container (only 2 GPUs were available when it was measured).
#pragma once
#include <stdint.h>

#include "machin_b200.h"

// ---- peer-memory exchange: pointer tables and the signal / wait halves of the barrier (used by the push kernels at the
// end of this file and by wire_compact_kernel)
struct PeerPtrs {
    char *p[MB_MAX_PEERS];
};

// Barrier folded into its neighbours (instead of a separate barrier launch between producer and consumer):
//   producer  every CTA fences its peer stores at system scope and takes a ticket; the LAST CTA of the grid publishes
//             epoch = state[0] + 1 into flag `self` of every member and stores it back into state[0]
//   consumer  every CTA waits until the P flags of THIS member reached state[0] (the local producer ran earlier on the
//             same stream, so state[0] already holds the current epoch), then reads what the members pushed.
// flags == nullptr switches both off (the NCCL transport, the emulation of the plain kernels).
struct PeerSignal {
    PeerPtrs flags;                 // member m's flag row of this site, as mapped here
    unsigned long long *state;      // [0] epoch, [1] timed-out
    unsigned int *ticket;           // zero between launches
    int P, self_block, on;
};

__device__ __forceinline__ void peer_signal_tail(const PeerSignal &sg, unsigned total_ctas) {
    if (!sg.on) return;
    __threadfence_system();                         // my CTA's peer stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(sg.ticket, 1u);
        if (t == total_ctas - 1) {                  // ... and so are everybody else's: they fenced before their ticket
            *sg.ticket = 0;
            __threadfence_system();
            const unsigned long long epoch = sg.state[0] + 1;
            for (int r = 0; r < sg.P; ++r)
                *(reinterpret_cast<volatile unsigned long long *>(sg.flags.p[r]) + sg.self_block) = epoch;
            sg.state[0] = epoch;
        }
    }
}

__device__ __forceinline__ void peer_wait_head(const PeerSignal &sg) {
    if (!sg.on) return;
    if ((int)threadIdx.x < sg.P) {
        const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(sg.state);
        const volatile unsigned long long *mine =
            reinterpret_cast<volatile unsigned long long *>(sg.flags.p[sg.self_block]) + threadIdx.x;
        long long spins = 0;
        while (*mine < epoch) {
            if (++spins > (1ll << 25)) {            // a member never arrived (~10 s of L2 polls): report, do not hang the GPU
                sg.state[1] = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
}


struct FieldList {
    mb_field f[MB_MAX_FIELDS];
    int n;
};

__device__ __forceinline__ void warp_copy(char *__restrict__ dst, const char *__restrict__ src, int64_t nbytes,
                                          int lane) {
    const uintptr_t mis = (uintptr_t)dst | (uintptr_t)src | (uintptr_t)nbytes;
    if ((mis & 15) == 0) {
        const uint4 *s = reinterpret_cast<const uint4 *>(src);
        uint4 *d = reinterpret_cast<uint4 *>(dst);
        for (int64_t k = lane; k < (nbytes >> 4); k += 32) d[k] = __ldcs(s + k);
    } else if ((mis & 7) == 0) {
        const uint2 *s = reinterpret_cast<const uint2 *>(src);
        uint2 *d = reinterpret_cast<uint2 *>(dst);
        for (int64_t k = lane; k < (nbytes >> 3); k += 32) d[k] = s[k];
    } else if ((mis & 3) == 0) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d = reinterpret_cast<uint32_t *>(dst);
        for (int64_t k = lane; k < (nbytes >> 2); k += 32) d[k] = s[k];
    } else {
        for (int64_t k = lane; k < nbytes; k += 32) dst[k] = src[k];
    }
}

// kGather: dense[i] <- ring[index[i]]     !kGather: ring[(first + i) % cap] <- dense[i]
template <bool kGather>
__global__ void __launch_bounds__(256) rows_move_kernel(char *ring, int64_t pitch, int64_t capacity,
                                                        const int64_t *__restrict__ index, int64_t first_slot,
                                                        int64_t n, FieldList fl) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        int64_t slot;
        if (kGather) {
            slot = index ? index[i] : i;  // NULL index: identity (unpacking rows received from a peer shard)
            slot = slot < 0 ? 0 : (slot >= capacity ? capacity - 1 : slot);  // never read outside the ring
        } else {
            slot = (first_slot + i) % capacity;
        }
        char *row = ring + slot * pitch;
        for (int f = 0; f < fl.n; ++f) {
            char *dense = (char *)fl.f[f].d_dense + i * fl.f[f].dense_pitch;
            if (kGather) warp_copy(dense, row + fl.f[f].row_offset, fl.f[f].nbytes, lane);
            else warp_copy(row + fl.f[f].row_offset, dense, fl.f[f].nbytes, lane);
        }
    }
}

// Shard-exchange metadata: every wire row carries [index i64 | version i64 | value f64] behind its payload
// (prioritized_buffer_d.py:293-303 returns index, version and is_weight next to the batch; :184-202 sends
// index, version and priority back).  One launch packs / unpacks all three columns.
__global__ void __launch_bounds__(256) rows_pack_meta_kernel(char *__restrict__ wire, int64_t width, int64_t meta_off,
                                                             int64_t n, const int64_t *__restrict__ index,
                                                             const int64_t *__restrict__ version,
                                                             const int64_t *__restrict__ version_table,
                                                             const double *__restrict__ v64,
                                                             const float *__restrict__ v32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char *m = wire + i * width + meta_off;
    const int64_t k = index[i];
    reinterpret_cast<int64_t *>(m)[0] = k;
    reinterpret_cast<int64_t *>(m)[1] = version ? version[i] : version_table[k];
    reinterpret_cast<double *>(m)[2] = v64 ? v64[i] : (double)v32[i];
}

__global__ void __launch_bounds__(256) rows_unpack_meta_kernel(const char *__restrict__ wire, int64_t width,
                                                               int64_t meta_off, int64_t n, int64_t *__restrict__ index,
                                                               int64_t *__restrict__ version, double *__restrict__ v64,
                                                               float *__restrict__ v32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char *m = wire + i * width + meta_off;
    index[i] = reinterpret_cast<const int64_t *>(m)[0];
    version[i] = reinterpret_cast<const int64_t *>(m)[1];
    const double v = reinterpret_cast<const double *>(m)[2];
    if (v64) v64[i] = v;
    if (v32) v32[i] = (float)v;
}

// ---- fixed-shape exchange helpers (counts are read on the device: graph-replayable) --------------------
__global__ void __launch_bounds__(256) wire_compact_kernel(const char *__restrict__ recv, int P, int64_t cap,
                                                           int64_t width, int64_t meta_off,
                                                           const int64_t *__restrict__ counts, char *__restrict__ merged,
                                                           int64_t merged_rows, int64_t *__restrict__ index,
                                                           int64_t *__restrict__ version, double *__restrict__ isw,
                                                           const double *__restrict__ scale, PeerSignal wait) {
    peer_wait_head(wait);            // (peer transport: every member's rows have landed; no-op otherwise)
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // merged row handled by this warp
    if (j >= merged_rows) return;
    int64_t base = 0;
    int m = 0;
    while (m < P && j >= base + counts[m]) base += counts[m++];
    char *dst = merged + j * width;
    if (m == P) {                                  // padding row: zero payload, weight 0
        for (int64_t k = lane * 8; k < width; k += 256) *reinterpret_cast<uint2 *>(dst + k) = make_uint2(0, 0);
        if (lane == 0) { index[j] = 0; version[j] = -1; isw[j] = 0.0; }
        return;
    }
    const char *src = recv + ((int64_t)m * cap + (j - base)) * width;
    // wire rows are payload (multiple of 16) + 24 bytes of metadata: 8-byte granularity
    for (int64_t k = lane * 8; k < width; k += 256) *reinterpret_cast<uint2 *>(dst + k) = __ldcs(reinterpret_cast<const uint2 *>(src + k));
    if (lane == 0) {
        const char *meta = src + meta_off;
        index[j] = reinterpret_cast<const int64_t *>(meta)[0];
        version[j] = reinterpret_cast<const int64_t *>(meta)[1];
        isw[j] = reinterpret_cast<const double *>(meta)[2] * (scale ? *scale : 1.0);
    }
}

// Compaction straight into the dense field tensors (compact + unpack in one launch): merged row j is split into its fields
// like rows_move_kernel<true> does, padding rows are zero-filled, metadata goes out as in wire_compact_kernel.
__global__ void __launch_bounds__(256) wire_compact_fields_kernel(const char *__restrict__ recv, int P, int64_t cap,
                                                                  int64_t width, int64_t meta_off,
                                                                  const int64_t *__restrict__ counts, FieldList fl,
                                                                  int64_t merged_rows, int64_t *__restrict__ index,
                                                                  int64_t *__restrict__ version, double *__restrict__ isw,
                                                                  const double *__restrict__ scale, PeerSignal wait) {
    peer_wait_head(wait);
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // merged row handled by this warp
    if (j >= merged_rows) return;
    int64_t base = 0;
    int m = 0;
    while (m < P && j >= base + counts[m]) base += counts[m++];
    if (m == P) {                                  // padding row: zero fields, weight 0
        for (int f = 0; f < fl.n; ++f) {
            char *dense = (char *)fl.f[f].d_dense + j * fl.f[f].dense_pitch;
            for (int64_t k = lane; k < fl.f[f].nbytes; k += 32) dense[k] = 0;
        }
        if (lane == 0) { index[j] = 0; version[j] = -1; isw[j] = 0.0; }
        return;
    }
    const char *src = recv + ((int64_t)m * cap + (j - base)) * width;
    for (int f = 0; f < fl.n; ++f)
        warp_copy((char *)fl.f[f].d_dense + j * fl.f[f].dense_pitch, src + fl.f[f].row_offset, fl.f[f].nbytes, lane);
    if (lane == 0) {
        const char *meta = src + meta_off;
        index[j] = reinterpret_cast<const int64_t *>(meta)[0];
        version[j] = reinterpret_cast<const int64_t *>(meta)[1];
        isw[j] = reinterpret_cast<const double *>(meta)[2] * (scale ? *scale : 1.0);
    }
}

__global__ void __launch_bounds__(256) wire_scatter_update_kernel(char *__restrict__ send, int P, int64_t cap,
                                                                  const int64_t *__restrict__ counts,
                                                                  const int64_t *__restrict__ index,
                                                                  const int64_t *__restrict__ version,
                                                                  const float *__restrict__ priority,
                                                                  int64_t merged_rows) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= merged_rows) return;
    int64_t base = 0;
    int m = 0;
    while (m < P && j >= base + counts[m]) base += counts[m++];
    if (m == P) return;
    char *dst = send + ((int64_t)m * cap + (j - base)) * 24;
    reinterpret_cast<int64_t *>(dst)[0] = index[j];
    reinterpret_cast<int64_t *>(dst)[1] = version[j];
    reinterpret_cast<double *>(dst)[2] = (double)priority[j];
}

__global__ void __launch_bounds__(256) wire_update_mask_kernel(const char *__restrict__ recv, int P, int64_t cap,
                                                               const int64_t *__restrict__ counts, int self_block,
                                                               const int64_t *__restrict__ table, int64_t table_size,
                                                               int64_t *__restrict__ index, float *__restrict__ priority,
                                                               uint8_t *__restrict__ valid) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)P * cap) return;
    const int64_t i = e % cap;
    const char *src = recv + e * 24;
    const int64_t k = reinterpret_cast<const int64_t *>(src)[0];
    const int64_t v = reinterpret_cast<const int64_t *>(src)[1];
    const bool real = i < counts[self_block] && k >= 0 && k < table_size;
    index[e] = real ? k : 0;
    priority[e] = real ? (float)reinterpret_cast<const double *>(src)[2] : 0.0f;
    valid[e] = (real && table[k] == v) ? 1 : 0;
}

// Apportioning of one sharded sample on the device (prioritized_buffer_d.py:229-234, :425-427), so that a
// fixed-shape update needs no host read of the shard sums: total = sum(shard sums) + 1e-6 (left to right, like
// Python's sum), counts[m] = ceil(sums[m] * batch / total), scale = merged_rows / sum(counts) (padding rows weigh
// 0, so the mean over merged_rows rows times scale is the mean over the real ones), and one mb_per_params block
// per requester this shard serves: beta advances by beta_increment per draw (only when this shard draws at all),
// the final beta goes back to *beta.  One thread: P is the number of GPUs.
__device__ __forceinline__ void wire_apportion_body(const double *__restrict__ sums, int P, int64_t batch, int self_block,
                                                    int64_t merged_rows, const int64_t *__restrict__ n_stored, double *beta,
                                                    double beta_increment, int64_t *__restrict__ counts,
                                                    double *__restrict__ scale, mb_per_params *__restrict__ params) {
    double total = 0.0;
    for (int m = 0; m < P; ++m) total = __dadd_rn(total, sums[m]);
    total = __dadd_rn(total, 1e-6);
    int64_t n = 0;
    for (int m = 0; m < P; ++m) {
        const int64_t c = (int64_t)ceil(__ddiv_rn(__dmul_rn(sums[m], (double)batch), total));
        counts[m] = c;
        n += c;
    }
    *scale = n > 0 ? __ddiv_rn((double)merged_rows, (double)n) : 0.0;
    const int64_t mine = counts[self_block];
    double b = *beta;
    for (int r = 0; r < P; ++r) {
        params[r].beta = b;
        params[r].all_weight_sum = total;
        params[r].n_stored = *n_stored;
        params[r].step = 0;
        params[r].n_draw = mine;
        if (mine > 0) b = fmin(1.0, __dadd_rn(b, beta_increment));
    }
    *beta = b;
}

__global__ void wire_apportion_kernel(const double *__restrict__ sums, int P, int64_t batch, int self_block,
                                      int64_t merged_rows, const int64_t *__restrict__ n_stored, double *beta,
                                      double beta_increment, int64_t *__restrict__ counts, double *__restrict__ scale,
                                      mb_per_params *__restrict__ params) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    wire_apportion_body(sums, P, batch, self_block, merged_rows, n_stored, beta, beta_increment, counts, scale, params);
}


// ---- shard exchange over peer memory (NVLink): the same data movement as gather + pack_meta + all-to-all, and as
// scatter_update + all-to-all, with the destination in ANOTHER GPU's buffer (peer-mapped pointers).  Plain stores:
// the ordering against the consumer is the peer barrier that follows (csrc/peer.cu).
__global__ void __launch_bounds__(256) wire_push_rows_kernel(const char *__restrict__ ring, int64_t pitch,
                                                             int64_t capacity, const int64_t *__restrict__ index,
                                                             const double *__restrict__ isw,
                                                             const int64_t *__restrict__ version_table, int P,
                                                             int64_t cap, int self_block, int64_t width, PeerPtrs recv,
                                                             PeerSignal sg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < (int64_t)P * cap; j += n_warps) {
        const int r = (int)(j / cap);                      // requester this row was drawn for
        const int64_t i = j - (int64_t)r * cap;
        int64_t slot = index[j];
        slot = slot < 0 ? 0 : (slot >= capacity ? capacity - 1 : slot);
        char *dst = recv.p[r] + ((int64_t)self_block * cap + i) * width;
        warp_copy(dst, ring + slot * pitch, pitch, lane);
        if (lane == 0) {
            char *m = dst + pitch;
            reinterpret_cast<int64_t *>(m)[0] = index[j];
            reinterpret_cast<int64_t *>(m)[1] = version_table[slot];
            reinterpret_cast<double *>(m)[2] = isw[j];
        }
    }
    peer_signal_tail(sg, gridDim.x);
}

__global__ void __launch_bounds__(256) wire_push_update_kernel(int P, int64_t cap, int self_block,
                                                               const int64_t *__restrict__ counts,
                                                               const int64_t *__restrict__ index,
                                                               const int64_t *__restrict__ version,
                                                               const float *__restrict__ priority, int64_t merged_rows,
                                                               PeerPtrs recv) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= merged_rows) return;
    int64_t base = 0;
    int m = 0;
    while (m < P && j >= base + counts[m]) base += counts[m++];
    if (m == P) return;                                    // padding row of the merged batch
    char *dst = recv.p[m] + ((int64_t)self_block * cap + (j - base)) * 24;
    reinterpret_cast<int64_t *>(dst)[0] = index[j];
    reinterpret_cast<int64_t *>(dst)[1] = version[j];
    reinterpret_cast<double *>(dst)[2] = (double)priority[j];
}

// nbytes of src -> slot `self_block` of every member (blockIdx.y = destination member)
__global__ void __launch_bounds__(256) peer_push_kernel(const char *__restrict__ src, int64_t nbytes, int self_block,
                                                        int64_t slot_stride, PeerPtrs dst, PeerSignal sg) {
    char *out = dst.p[blockIdx.y] + (int64_t)self_block * slot_stride;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)src | (uintptr_t)out | (uintptr_t)nbytes) & 15) == 0) {
        for (int64_t k = tid; k < (nbytes >> 4); k += stride)
            reinterpret_cast<uint4 *>(out)[k] = reinterpret_cast<const uint4 *>(src)[k];
    } else if ((((uintptr_t)src | (uintptr_t)out | (uintptr_t)nbytes) & 7) == 0) {
        for (int64_t k = tid; k < (nbytes >> 3); k += stride)
            reinterpret_cast<uint2 *>(out)[k] = reinterpret_cast<const uint2 *>(src)[k];
    } else {
        for (int64_t k = tid; k < nbytes; k += stride) out[k] = src[k];
    }
    peer_signal_tail(sg, gridDim.x * gridDim.y);
}

// out[i] = (slots[0][i] + slots[1][i] + ... + slots[P-1][i]) / divisor, member order (separately rounded adds)
__global__ void __launch_bounds__(256) peer_sum_kernel(const float *__restrict__ slots, int P, int64_t n, int64_t slot_floats,
                                                       float divisor, float *__restrict__ out, PeerSignal wait) {
    peer_wait_head(wait);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = slots[i];
        for (int m = 1; m < P; ++m) acc = __fadd_rn(acc, slots[(int64_t)m * slot_floats + i]);
        out[i] = __fdiv_rn(acc, divisor);
    }
}
