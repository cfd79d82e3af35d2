// Replay-row movement for sm_100a: batch assembly (gather) and ring insert (scatter).
//
// Replaces the reference's `[storage[i] for i in index]` + per-attribute torch.cat
// (prioritized_buffer.py:390-394, buffer.py:261-353,380-413) and the per-transition deepcopy ring
// insert (storage.py:85-114).  The device ring is an array of packed rows (`pitch` bytes, multiple of
// 16); one warp moves one row with 16-byte accesses wherever offsets and pitches allow it, so a
// sampled transition costs one contiguous HBM segment instead of one Python object per attribute.
#include <stdlib.h>

#include "mb_common.cuh"

namespace {

#include "rows_device.cuh"

#include "rows_tma.cuh"   // rows_gather_bulk_kernel (cp.async.bulk / mbarrier)

// MB_ROWS_LEGACY=1: the one-warp-per-row gather everywhere (A/B timing and the parity test of the two kernels)
bool rows_legacy() {
    const char *e = getenv("MB_ROWS_LEGACY");
    return e && e[0] == '1';
}

// Bulk-copy gather if the wide fields qualify: every field whose offset / size / destination / pitch are multiples
// of 16 leaves shared memory with one cp.async.bulk; it pays when those fields carry most of the bytes.
bool launch_gather_bulk(const char *ring, int64_t pitch, int64_t capacity, const int64_t *index, int64_t n,
                        const FieldList &fl, cudaStream_t s) {
    if (rows_legacy() || (pitch & 15) || ((uintptr_t)ring & 15) || pitch > 48 * 1024) return false;
    unsigned mask = 0;
    int64_t wide = 0, all = 0;
    for (int f = 0; f < fl.n; ++f) {
        const mb_field &x = fl.f[f];
        all += x.nbytes;
        if (x.nbytes >= 16 && ((x.row_offset | x.nbytes | x.dense_pitch | (int64_t)(uintptr_t)x.d_dense) & 15) == 0) {
            mask |= 1u << f;
            wide += x.nbytes;
        }
    }
    if (2 * wide < all) return false;
    // rows per tile: enough tiles to cover every SM twice at small n, at most 16 rows and 12 KB per stage
    int64_t R = n / (2 * mb::kNumSMs);
    R = R < 1 ? 1 : (R > 16 ? 16 : R);
    while (R > 1 && R * pitch > 12 * 1024) --R;
    const int64_t stage = (R * pitch + 127) & ~(int64_t)127;
    const size_t dyn = (size_t)(stage * kBulkStages);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(rows_gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set = true;
    }
    const int64_t n_tiles = (n + R - 1) / R;
    int64_t per_sm = (int64_t)(200 * 1024) / (int64_t)(dyn + 1024);
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    int64_t blocks = (int64_t)mb::kNumSMs * per_sm;
    if (blocks > n_tiles) blocks = n_tiles;
    MB_PREFER_MAX_SMEM(rows_gather_bulk_kernel);
    rows_gather_bulk_kernel<<<(unsigned)blocks, kBulkThreads, dyn, s>>>(ring, pitch, capacity, index, n, fl, (int)R, mask);
    return true;
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
    if (gather && launch_gather_bulk((const char *)ring, pitch, capacity, index, n, fl, s)) {
        MB_LAUNCHED(1);
        return MB_OK;
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

int mb_rows_pack_meta(void *d_wire, int64_t width, int64_t meta_offset, int64_t n, const int64_t *d_index,
                      const int64_t *d_version, const int64_t *d_version_table, const double *d_value_f64,
                      const float *d_value_f32, void *stream) {
    if (!d_wire || n < 0 || !d_index || (!d_version && !d_version_table) || (!d_value_f64 && !d_value_f32) ||
        meta_offset % 8 || width % 8 || meta_offset + 24 > width)
        return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    rows_pack_meta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (char *)d_wire, width, meta_offset, n, d_index, d_version, d_version_table, d_value_f64, d_value_f32);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_rows_unpack_meta(const void *d_wire, int64_t width, int64_t meta_offset, int64_t n, int64_t *d_index,
                        int64_t *d_version, double *d_value_f64, float *d_value_f32, void *stream) {
    if (!d_wire || n < 0 || !d_index || !d_version || meta_offset % 8 || width % 8 || meta_offset + 24 > width)
        return MB_ERR_INVALID;
    if (n == 0) return MB_OK;
    rows_unpack_meta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_wire, width, meta_offset, n, d_index, d_version, d_value_f64, d_value_f32);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_compact(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                    const int64_t *d_counts, void *d_merged, int64_t merged_rows, int64_t *d_index, int64_t *d_version,
                    double *d_is_weight, const double *d_scale, void *stream) {
    if (!d_recv || !d_counts || !d_merged || !d_index || !d_version || !d_is_weight || n_blocks < 1 || cap < 1 ||
        width % 8 || meta_offset % 8 || meta_offset + 24 > width || merged_rows < 0)
        return MB_ERR_INVALID;
    if (merged_rows == 0) return MB_OK;
    wire_compact_kernel<<<(unsigned)((merged_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_recv, n_blocks, cap, width, meta_offset, d_counts, (char *)d_merged, merged_rows, d_index,
        d_version, d_is_weight, d_scale, PeerSignal{});
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_scatter_update(void *d_send, int n_blocks, int64_t cap, const int64_t *d_counts, const int64_t *d_index,
                           const int64_t *d_version, const float *d_priority, int64_t merged_rows, void *stream) {
    if (!d_send || !d_counts || !d_index || !d_version || !d_priority || n_blocks < 1 || cap < 1 || merged_rows < 0)
        return MB_ERR_INVALID;
    if (merged_rows == 0) return MB_OK;
    wire_scatter_update_kernel<<<(unsigned)((merged_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (char *)d_send, n_blocks, cap, d_counts, d_index, d_version, d_priority, merged_rows);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_apportion(const double *d_sums, int n_blocks, int64_t batch, int self_block, int64_t merged_rows,
                      const int64_t *d_n_stored, double *d_beta, double beta_increment, int64_t *d_counts,
                      double *d_scale, mb_per_params *d_params, void *stream) {
    if (!d_sums || n_blocks < 1 || batch < 1 || self_block < 0 || self_block >= n_blocks || merged_rows < 1 ||
        !d_n_stored || !d_beta || !d_counts || !d_scale || !d_params)
        return MB_ERR_INVALID;
    wire_apportion_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_sums, n_blocks, batch, self_block, merged_rows,
                                                              d_n_stored, d_beta, beta_increment, d_counts, d_scale,
                                                              d_params);
    MB_LAUNCHED(1);
    return MB_OK;
}

int mb_wire_update_mask(const void *d_recv, int n_blocks, int64_t cap, const int64_t *d_counts, int self_block,
                        const int64_t *d_version_table, int64_t table_size, int64_t *d_index, float *d_priority,
                        uint8_t *d_valid, void *stream) {
    if (!d_recv || !d_counts || !d_version_table || !d_index || !d_priority || !d_valid || n_blocks < 1 || cap < 1 ||
        self_block < 0 || self_block >= n_blocks)
        return MB_ERR_INVALID;
    const int64_t total = (int64_t)n_blocks * cap;
    wire_update_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_recv, n_blocks, cap, d_counts, self_block, d_version_table, table_size, d_index, d_priority,
        d_valid);
    MB_LAUNCHED(1);
    return MB_OK;
}

}  // extern "C"
