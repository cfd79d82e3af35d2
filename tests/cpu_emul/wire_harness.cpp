// TEST INFRASTRUCTURE ONLY -- compiles machin_b200/csrc/rows_device.cuh (gather / scatter of packed rows, wire
// metadata, the fixed-shape exchange helpers) for the host through cuda_shim.h; entry points for
// tests/test_exchange_emulated_cpu.py mirror the launch geometry of csrc/rows.cu.
#include "cuda_shim.h"

struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
template <typename T> inline T __ldcs(const T *p) { return *p; }

#undef __shared__
#define __shared__
#include "rows_device.cuh"

extern "C" {

int emul_rows_gather(const char *ring, int64_t pitch, int64_t capacity, const int64_t *index, int64_t n,
                     const mb_field *fields, int n_fields) {
    FieldList fl{};
    fl.n = n_fields;
    for (int i = 0; i < n_fields; ++i) fl.f[i] = fields[i];
    int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 64) blocks = 64;
    shim_launch((unsigned)blocks, 256, [&] {
        rows_move_kernel<true>(const_cast<char *>(ring), pitch, capacity, index, 0, n, fl);
    });
    return 0;
}

int emul_pack_meta(char *wire, int64_t width, int64_t meta_off, int64_t n, const int64_t *index,
                   const int64_t *version_table, const double *v64) {
    shim_launch((unsigned)((n + 255) / 256), 256, [&] {
        rows_pack_meta_kernel(wire, width, meta_off, n, index, nullptr, version_table, v64, nullptr);
    });
    return 0;
}

int emul_wire_compact(const char *recv, int P, int64_t cap, int64_t width, int64_t meta_off, const int64_t *counts,
                      char *merged, int64_t merged_rows, int64_t *index, int64_t *version, double *isw,
                      const double *scale) {
    shim_launch((unsigned)((merged_rows * 32 + 255) / 256), 256, [&] {
        wire_compact_kernel(recv, P, cap, width, meta_off, counts, merged, merged_rows, index, version, isw, scale, PeerSignal{});
    });
    return 0;
}

int emul_wire_scatter_update(char *send, int P, int64_t cap, const int64_t *counts, const int64_t *index,
                             const int64_t *version, const float *priority, int64_t merged_rows) {
    shim_launch((unsigned)((merged_rows + 255) / 256), 256, [&] {
        wire_scatter_update_kernel(send, P, cap, counts, index, version, priority, merged_rows);
    });
    return 0;
}

int emul_wire_update_mask(const char *recv, int P, int64_t cap, const int64_t *counts, int self_block,
                          const int64_t *table, int64_t table_size, int64_t *index, float *priority, uint8_t *valid) {
    shim_launch((unsigned)(((int64_t)P * cap + 255) / 256), 256, [&] {
        wire_update_mask_kernel(recv, P, cap, counts, self_block, table, table_size, index, priority, valid);
    });
    return 0;
}

int emul_wire_apportion(const double *sums, int P, int64_t batch, int self_block, int64_t merged_rows,
                        const int64_t *n_stored, double *beta, double beta_increment, int64_t *counts, double *scale,
                        mb_per_params *params) {
    shim_launch(1, 32, [&] {
        wire_apportion_kernel(sums, P, batch, self_block, merged_rows, n_stored, beta, beta_increment, counts, scale, params);
    });
    return 0;
}

int emul_wire_compact_fields(const char *recv, int P, int64_t cap, int64_t width, int64_t meta_off, const int64_t *counts,
                             const mb_field *fields, int n_fields, int64_t merged_rows, int64_t *index, int64_t *version,
                             double *isw, const double *scale) {
    FieldList fl{};
    fl.n = n_fields;
    for (int i = 0; i < n_fields; ++i) fl.f[i] = fields[i];
    shim_launch((unsigned)((merged_rows * 32 + 255) / 256), 256, [&] {
        wire_compact_fields_kernel(recv, P, cap, width, meta_off, counts, fl, merged_rows, index, version, isw, scale, PeerSignal{});
    });
    return 0;
}

// peer-memory exchange: "peers" are plain host buffers here (one per emulated shard)
int emul_wire_push_rows(const char *ring, int64_t pitch, int64_t capacity, const int64_t *index, const double *isw,
                        const int64_t *version_table, int P, int64_t cap, int self_block, int64_t width, char **peers) {
    PeerPtrs pp{};
    for (int m = 0; m < P; ++m) pp.p[m] = peers[m];
    int64_t blocks = ((int64_t)P * cap * 32 + 255) / 256;
    if (blocks > 64) blocks = 64;
    shim_launch((unsigned)blocks, 256, [&] {
        wire_push_rows_kernel(ring, pitch, capacity, index, isw, version_table, P, cap, self_block, width, pp, PeerSignal{});
    });
    return 0;
}

int emul_wire_push_update(int P, int64_t cap, int self_block, const int64_t *counts, const int64_t *index,
                          const int64_t *version, const float *priority, int64_t merged_rows, char **peers) {
    PeerPtrs pp{};
    for (int m = 0; m < P; ++m) pp.p[m] = peers[m];
    shim_launch((unsigned)((merged_rows + 255) / 256), 256, [&] {
        wire_push_update_kernel(P, cap, self_block, counts, index, version, priority, merged_rows, pp);
    });
    return 0;
}

int emul_peer_push(const char *src, int64_t nbytes, int P, int self_block, int64_t slot_stride, char **peers) {
    PeerPtrs pp{};
    for (int m = 0; m < P; ++m) pp.p[m] = peers[m];
    for (int dst = 0; dst < P; ++dst) {                 // blockIdx.y = destination member
        blockIdx.y = (unsigned)dst;
        shim_launch(2, 256, [&] { blockIdx.y = (unsigned)dst; peer_push_kernel(src, nbytes, self_block, slot_stride, pp, PeerSignal{}); });
    }
    return 0;
}

int emul_peer_sum(const float *slots, int P, int64_t n, float divisor, float *out) {
    shim_launch(2, 256, [&] { peer_sum_kernel(slots, P, n, n, divisor, out, PeerSignal{}); });
    return 0;
}

}  // extern "C"
