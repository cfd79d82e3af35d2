// Batch assembly with the bulk-copy engine (TMA, `cp.async.bulk`, sm_90+; SASS: UBLKCP): device code for rows.cu only
// (inline PTX -- not part of the g++ emulation harnesses, which keep covering rows_move_kernel).
//
// A sampled transition is ONE contiguous, 16-byte aligned HBM segment (`pitch` bytes).  rows_move_kernel moves it
// with one warp per row: index load -> 16-byte loads -> stores, i.e. two dependent HBM latencies per warp and at most
// 688 bytes in flight per warp.  Here a CTA owns a tile of R rows per pipeline stage:
//   producer  warp 0 reads the R indexes and issues ONE bulk copy per row, HBM -> shared memory, all completing on the
//             stage's mbarrier (expect_tx = R * pitch): the copy engine generates the addresses, the SM only issues
//   consumer  after the mbarrier flips, every wide field of every row (16-byte aligned offset / size / destination:
//             state, next_state) leaves with ONE bulk copy shared memory -> HBM into its dense [B, ...] tensor; the
//             narrow fields (action, reward, terminal: 8 / 4 / 1 bytes) are stored by single threads
//   kStages stages are in flight per CTA; CTAs are persistent (grid = a multiple of the SM count).
// Same bytes, same destinations as rows_move_kernel<true>: the parity tests run both.
#pragma once
#include <stdint.h>

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma

constexpr int kBulkStages = 4;
constexpr int kBulkThreads = 128;

// bulk_mask bit f: field f can leave shared memory with a bulk copy (offset, size, destination and pitch all
// multiples of 16 -- checked on the host).  index == nullptr: identity (unpacking rows received from a peer).
__global__ void __launch_bounds__(kBulkThreads) rows_gather_bulk_kernel(const char *__restrict__ ring, int64_t pitch,
                                                                        int64_t capacity,
                                                                        const int64_t *__restrict__ index, int64_t n,
                                                                        FieldList fl, int rows_per_tile,
                                                                        unsigned bulk_mask) {
    MB_DYNAMIC_SMEM(dyn);                      // kBulkStages x rows_per_tile x pitch, 128-byte aligned stages
    __shared__ __align__(8) uint64_t full[kBulkStages];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int R = rows_per_tile;
    const int64_t stage_bytes = ((int64_t)R * pitch + 127) & ~(int64_t)127;
    const int64_t n_tiles = (n + R - 1) / R;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kBulkStages; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int64_t k) {              // warp 0, all lanes: tile k of this CTA -> stage k % kBulkStages
        if (k >= my_tiles) return;
        const int s = (int)(k % kBulkStages);
        const int64_t first = ((int64_t)blockIdx.x + k * gridDim.x) * R;
        const int rows = (int)((n - first) < R ? (n - first) : R);
        if (lane == 0) tma::mbar_expect_tx(&full[s], (uint32_t)(rows * pitch));
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
            int64_t slot = index ? index[first + r] : first + r;
            slot = slot < 0 ? 0 : (slot >= capacity ? capacity - 1 : slot);   // never read outside the ring
            tma::bulk_g2s(dyn + s * stage_bytes + (int64_t)r * pitch, ring + slot * pitch, (uint32_t)pitch, &full[s]);
        }
    };
    if (wid == 0)
        for (int k = 0; k < kBulkStages; ++k) issue(k);
    for (int64_t k = 0; k < my_tiles; ++k) {
        const int s = (int)(k % kBulkStages);
        tma::mbar_wait(&full[s], (uint32_t)((k / kBulkStages) & 1));
        const int64_t first = ((int64_t)blockIdx.x + k * gridDim.x) * R;
        const int rows = (int)((n - first) < R ? (n - first) : R);
        const unsigned char *base = dyn + s * stage_bytes;
        for (int item = tid; item < rows * fl.n; item += kBulkThreads) {
            const int f = item / rows, r = item - f * rows;                  // field-major: a warp mostly works on one field
            const unsigned char *src = base + (int64_t)r * pitch + fl.f[f].row_offset;
            char *dst = (char *)fl.f[f].d_dense + (first + r) * fl.f[f].dense_pitch;
            const int64_t nb = fl.f[f].nbytes;
            if ((bulk_mask >> f) & 1u) {
                tma::bulk_s2g(dst, src, (uint32_t)nb);
            } else if (((nb | fl.f[f].row_offset | fl.f[f].dense_pitch | (int64_t)(uintptr_t)fl.f[f].d_dense) & 7) == 0) {
                for (int64_t b = 0; b < nb; b += 8) *reinterpret_cast<uint64_t *>(dst + b) = *reinterpret_cast<const uint64_t *>(src + b);
            } else if (((nb | fl.f[f].row_offset | fl.f[f].dense_pitch | (int64_t)(uintptr_t)fl.f[f].d_dense) & 3) == 0) {
                for (int64_t b = 0; b < nb; b += 4) *reinterpret_cast<uint32_t *>(dst + b) = *reinterpret_cast<const uint32_t *>(src + b);
            } else {
                for (int64_t b = 0; b < nb; ++b) dst[b] = (char)src[b];
            }
        }
        tma::bulk_commit();
        tma::bulk_wait_read0();                // my bulk stores have finished READING this stage
        __syncthreads();                       // ... and so have everybody else's: the stage can be refilled
        if (wid == 0) issue(k + kBulkStages);
    }
    tma::bulk_wait0();                         // all my bulk stores have landed before the CTA retires
}
