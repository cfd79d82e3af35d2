// Replay-row movement for sm_100a: batch assembly (gather) and ring insert (scatter).
//
// Replaces the reference's `[storage[i] for i in index]` + per-attribute torch.cat
// (prioritized_buffer.py:390-394, buffer.py:261-353,380-413) and the per-transition deepcopy ring
// insert (storage.py:85-114).  The device ring is an array of packed rows (`pitch` bytes, multiple of
// 16); one warp moves one row with 16-byte accesses wherever offsets and pitches allow it, so a
// sampled transition costs one contiguous HBM segment instead of one Python object per attribute.
#include "mb_common.cuh"

namespace {

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

int launch_rows(bool gather, void *ring, int64_t pitch, int64_t capacity, const int64_t *index,
                int64_t first_slot, int64_t n, const mb_field *fields, int n_fields, cudaStream_t s) {
    if (!ring || pitch <= 0 || capacity <= 0 || n < 0 || n_fields < 0 || n_fields > MB_MAX_FIELDS ||
        (n_fields > 0 && !fields))
        return MB_ERR_INVALID;
    if (n == 0 || n_fields == 0) return MB_OK;
    FieldList fl;
    fl.n = n_fields;
    for (int i = 0; i < n_fields; ++i) {
        fl.f[i] = fields[i];
        if (fields[i].row_offset < 0 || fields[i].nbytes < 0 || fields[i].row_offset + fields[i].nbytes > pitch)
            return MB_ERR_INVALID;
    }
    int64_t blocks = (n + 7) / 8;  // 8 warps (rows) per CTA
    const int64_t cap = (int64_t)mb::kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (gather)
        rows_move_kernel<true><<<(unsigned)blocks, 256, 0, s>>>((char *)ring, pitch, capacity, index, 0, n, fl);
    else
        rows_move_kernel<false><<<(unsigned)blocks, 256, 0, s>>>((char *)ring, pitch, capacity, nullptr, first_slot, n, fl);
    MB_LAUNCHED(1);
    return MB_OK;
}

}  // namespace

extern "C" {

int mb_rows_gather(const void *d_ring, int64_t pitch, int64_t capacity, const int64_t *d_index, int64_t n,
                   const mb_field *fields, int n_fields, void *stream) {
    return launch_rows(true, (void *)d_ring, pitch, capacity, d_index, 0, n, fields, n_fields, (cudaStream_t)stream);
}

int mb_rows_scatter(void *d_ring, int64_t pitch, int64_t capacity, int64_t first_slot, int64_t n,
                    const mb_field *fields, int n_fields, void *stream) {
    if (first_slot < 0 || n > capacity) return MB_ERR_INVALID;
    return launch_rows(false, d_ring, pitch, capacity, nullptr, first_slot, n, fields, n_fields, (cudaStream_t)stream);
}

int mb_rows_gather_packed(const void *d_ring, int64_t pitch, int64_t capacity, const int64_t *d_index, int64_t n,
                          void *d_out, void *stream) {
    mb_field f{0, pitch, d_out, pitch};
    return mb_rows_gather(d_ring, pitch, capacity, d_index, n, &f, 1, stream);
}

}  // extern "C"
