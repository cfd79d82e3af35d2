/*
 * machin_b200.h -- C ABI of libmachin_b200.so: the B200 (sm_100a) learner hot path of
 * iffiX/machin's massively parallel algorithms (APEX-DQN / APEX-DDPG / IMPALA / A3C).
 *
 * The reference has no FFI: the path lives behind Python classes.
 * Each entry point below is what a binding for that class/method would call; the
 * reference interface it replaces is cited as file:line relative to the reference root.
 * INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every function returns an int status: MB_OK (0) or a negative mb_status
 *   - pointers named d_* are DEVICE pointers (on the device the handle / call names),
 *     pointers named h_* are HOST pointers
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as void*;
 *     NULL = the legacy default stream) and never synchronise unless stated
 *   - *_host entry points take host buffers, stage them through pinned memory, run the same
 *     kernels and block until the results are in the caller's host buffers
 *   - no internal locking: callers serialise access to one handle (the reference keeps an
 *     RLock around buffer mutation, prioritized_buffer_d.py:109)
 */
#ifndef MACHIN_B200_H
#define MACHIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 1

typedef enum mb_status {
    MB_OK = 0,
    MB_ERR_CUDA = -1,          /* a CUDA runtime call failed; see mb_last_cuda_error()             */
    MB_ERR_INDEX_OOB = -2,     /* ValueError("Index has elements out of boundary!")   :90,136,173  */
    MB_ERR_DIM_MISMATCH = -3,  /* ValueError("Dimension of weights and indexes doesn't match!") :165 */
    MB_ERR_SIZE_MISMATCH = -4, /* ValueError("Weights size must match tree size!")    :197         */
    MB_ERR_INVALID = -5,       /* bad argument (NULL handle, negative count, unsupported dtype)     */
    MB_ERR_NOMEM = -6,         /* device or pinned allocation failed                               */
    MB_ERR_NO_DEVICE = -7      /* no CUDA device: the product path never falls back to the CPU      */
} mb_status;

typedef enum mb_dtype { MB_F32 = 0, MB_F64 = 1 } mb_dtype;

int mb_abi_version(void);
const char *mb_status_string(int status);
const char *mb_last_cuda_error(void);
int mb_device_count(void);
/* number of CUDA kernels this library has launched in this process (monotonic; bench.py reports
 * the difference over its timed region as gpu_launches) */
long long mb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Sum tree  (machin/frame/buffers/prioritized_buffer.py:8-231, class WeightTree)
 *
 * Device-resident float64 tree in the reference's flattened layout: level 0 (leaves, padded
 * to a power of two) first, root last; depth = ceil(log2(size)) + 1; 2^depth - 1 nodes.
 * ------------------------------------------------------------------------------------------ */
typedef struct mb_tree mb_tree;

/* WeightTree.__init__ (:13-52).  Allocates and zeroes the tree on `device`. */
int mb_tree_create(int64_t size, int device, mb_tree **out);
int mb_tree_destroy(mb_tree *tree);
/* depth / padded leaf count / node count of the flattened array (attributes depth, sizes, offsets). */
int mb_tree_geometry(const mb_tree *tree, int64_t *size, int *depth, int64_t *n_leaf_padded,
                     int64_t *n_nodes);
/* device pointer to the flattened float64 node array (attribute `weights`, :52) */
double *mb_tree_weights(mb_tree *tree);
/* get_weight_sum / get_leaf_max (:54-66).  Blocks on `stream`, then returns the scalars.
 * h_max_leaf_dtype (nullable) receives MB_F32 when the running maximum was set by a float32
 * batch: the reference then holds it as np.float32, which decides the dtype in which
 * normalize(max_leaf) is evaluated for default-priority stores (:302-303). */
int mb_tree_stats(mb_tree *tree, double *h_weight_sum, double *h_max_leaf, int *h_max_leaf_dtype,
                  void *stream);
/* device pointers to the root node and to the running leaf maximum (no sync) */
const double *mb_tree_root_ptr(mb_tree *tree);
const double *mb_tree_max_leaf_ptr(mb_tree *tree);
/* device address + size of the tree's scalar state (running max and its dtype flag, deferred
 * error): with mb_tree_weights() this is everything needed to snapshot / restore a tree in place */
void *mb_tree_state_blob(mb_tree *tree, int64_t *nbytes);
/* Deferred error of asynchronous device-side index validation: returns MB_OK or
 * MB_ERR_INDEX_OOB (and clears it).  Blocks on `stream`. */
int mb_tree_check(mb_tree *tree, void *stream);
/* zero all nodes, max_leaf = 0 (PrioritizedBuffer.clear builds a fresh tree, :311-317) */
int mb_tree_clear(mb_tree *tree, void *stream);

/* update_all_leaves + _build (:188-199, :216-231).  d_leaves: f64[n], n must equal size
 * (else MB_ERR_SIZE_MISMATCH).  d_leaves == NULL rebuilds from the leaves already in place. */
int mb_tree_build(mb_tree *tree, const double *d_leaves, int64_t n, void *stream);
/* find_leaf_index (:96-125).  d_query f64[n] -> d_index i64[n]. */
int mb_tree_find(mb_tree *tree, const double *d_query, int64_t n, int64_t *d_index, void *stream);
/* get_leaf_weight (:76-94).  d_index i64[n] -> d_weight f64[n]; out-of-range indexes raise the
 * deferred MB_ERR_INDEX_OOB (see mb_tree_check) and read as 0. */
int mb_tree_leaf_weight(mb_tree *tree, const int64_t *d_index, int64_t n, double *d_weight,
                        void *stream);
/* update_leaf_batch (:152-186).  d_weight is f32[n] or f64[n] (weight_dtype); duplicates in
 * d_index resolve last-write-wins like NumPy fancy assignment (:178); parents are recomputed
 * as left + right (:184-186), so the node array is bit-identical to the reference's.
 * If any index is out of [0, size) nothing is written and the deferred error is raised.
 * d_valid (nullable, u8[n]): entries with 0 are skipped -- the version filter of
 * DistributedPrioritizedBuffer._update_priority_service (prioritized_buffer_d.py:282-291). */
int mb_tree_update_batch(mb_tree *tree, const int64_t *d_index, const void *d_weight,
                         int weight_dtype, const uint8_t *d_valid, int64_t n, void *stream);
/* PrioritizedBuffer.update_priority (:319-328): w = (|p| + epsilon) ** alpha evaluated in the
 * dtype of p (_normalize_priority, :430-434), then update_leaf_batch.  */
int mb_tree_update_priority(mb_tree *tree, const int64_t *d_index, const void *d_priority,
                            int priority_dtype, const uint8_t *d_valid, int64_t n, double epsilon,
                            double alpha, void *stream);
/* update_leaf (:127-150): one leaf, path to root.  Host scalars; asynchronous. */
int mb_tree_update_leaf(mb_tree *tree, double weight, int64_t index, void *stream);

/* PrioritizedBuffer.sample_index_and_weight (:397-428), fused: stratified query from the
 * caller's uniforms (the reference draws them with np.random.uniform, :411), descent, leaf
 * weight, importance weight (n_stored * p / total) ** -beta, max-normalised.
 *   d_uniform   f64[n] in [0,1)        d_index i64[n]      d_is_weight f64[n]
 *   all_weight_sum <= 0 means "use the local root" (`all_weight_sum or ...`, :421)
 *   d_params (nullable) overrides the three scalars from device memory so that the call can be
 *   replayed inside a CUDA graph while beta / n_stored change. */
typedef struct mb_per_params {
    double beta;
    double all_weight_sum;
    int64_t n_stored;
    int64_t step;   /* which row of a [steps, n] uniform table to read (graph replay), else 0 */
    int64_t n_draw; /* > 0: draw this many samples (<= the n of the call, which then only sizes the launch):
                     * the shard exchange replays a fixed-shape graph while the apportioned count changes */
} mb_per_params;
int mb_per_sample(mb_tree *tree, const double *d_uniform, int64_t n, int64_t n_stored,
                  double all_weight_sum, double beta, const mb_per_params *d_params,
                  int64_t *d_index, double *d_is_weight, void *stream);

/* The same draw for n_requesters requesters in one launch (a shard serving every requester of a fixed-shape
 * update): requester r reads d_uniform[r * n ..], d_params[r] (device-side parameters are mandatory here) and
 * writes d_index[r * n ..] / d_is_weight[r * n ..].  Each requester's IS weights are normalised by its own max. */
int mb_per_sample_multi(mb_tree *tree, const double *d_uniform, int64_t n, int n_requesters,
                        const mb_per_params *d_params, int64_t *d_index, double *d_is_weight, void *stream);

/* Host-buffer variants (same kernels; H2D + launch + D2H + synchronise inside the call). */
int mb_tree_build_host(mb_tree *tree, const double *h_leaves, int64_t n);
int mb_tree_find_host(mb_tree *tree, const double *h_query, int64_t n, int64_t *h_index);
int mb_tree_leaf_weight_host(mb_tree *tree, const int64_t *h_index, int64_t n, double *h_weight);
int mb_tree_update_batch_host(mb_tree *tree, const int64_t *h_index, const void *h_weight,
                              int weight_dtype, int64_t n);
int mb_tree_read_host(mb_tree *tree, int64_t first_node, int64_t n_nodes, double *h_out);
int mb_per_sample_host(mb_tree *tree, const double *h_uniform, int64_t n, int64_t n_stored,
                       double all_weight_sum, double beta, int64_t *h_index, double *h_is_weight);

/* ------------------------------------------------------------------------------------------
 * Replay rows  (batch assembly: prioritized_buffer.py:390-394 + buffer.py:261-353,380-413;
 *               ring insert: storage.py:85-114)
 *
 * The device ring is an array of packed rows, `pitch` bytes apart (multiple of 16).  A field is
 * one attribute of the transition (state["k"], action["k"], reward, terminal, ...).
 * ------------------------------------------------------------------------------------------ */
#define MB_MAX_FIELDS 16
typedef struct mb_field {
    int64_t row_offset; /* byte offset of the field inside a packed row                       */
    int64_t nbytes;     /* bytes of the field per transition                                   */
    void *d_dense;      /* dense [n, nbytes] tensor: gather destination / scatter source       */
    int64_t dense_pitch;/* bytes between consecutive rows of d_dense (normally == nbytes)      */
} mb_field;

/* dense[i, :] = ring[index[i], field]   for every field ("[storage[i] for i in index]" + t.cat).
 * d_index == NULL means the identity (unpack n packed rows, e.g. rows received from a peer shard). */
int mb_rows_gather(const void *d_ring, int64_t pitch, int64_t capacity, const int64_t *d_index,
                   int64_t n, const mb_field *fields, int n_fields, void *stream);
/* ring[(first_slot + i) % capacity, field] = dense[i, :]   (TransitionStorageBasic.store_episode) */
int mb_rows_scatter(void *d_ring, int64_t pitch, int64_t capacity, int64_t first_slot, int64_t n,
                    const mb_field *fields, int n_fields, void *stream);
/* whole packed rows: out[i, :pitch] = ring[index[i], :]  (wire format of the shard exchange) */
int mb_rows_gather_packed(const void *d_ring, int64_t pitch, int64_t capacity,
                          const int64_t *d_index, int64_t n, void *d_out, void *stream);

/* Shard-exchange wire rows: [payload | index i64 | version i64 | value f64], `width` bytes apart, the three
 * metadata columns starting at `meta_offset`.  pack: version comes from d_version or is looked up as
 * d_version_table[index] (_sample_service, prioritized_buffer_d.py:300); value is the IS weight (f64) or a
 * priority (f32, widened exactly).  unpack: the inverse on the receiving member. */
int mb_rows_pack_meta(void *d_wire, int64_t width, int64_t meta_offset, int64_t n, const int64_t *d_index,
                      const int64_t *d_version, const int64_t *d_version_table, const double *d_value_f64,
                      const float *d_value_f32, void *stream);
int mb_rows_unpack_meta(const void *d_wire, int64_t width, int64_t meta_offset, int64_t n, int64_t *d_index,
                        int64_t *d_version, double *d_value_f64, float *d_value_f32, void *stream);
/* Device-side apportioning of one sharded sample (prioritized_buffer_d.py:229-234): from the all-gathered shard
 * sums d_sums[n_blocks] computes total = sum + 1e-6, d_counts[m] = ceil(sums[m] * batch / total),
 * *d_scale = merged_rows / sum(counts), and one mb_per_params block per requester for this shard's draws (beta from
 * *d_beta, advanced by beta_increment per draw and written back; n_stored from *d_n_stored; n_draw = counts[self]).
 * Lets a fixed-shape update replay without the host ever reading the shard sums. */
int mb_wire_apportion(const double *d_sums, int n_blocks, int64_t batch, int self_block, int64_t merged_rows,
                      const int64_t *d_n_stored, double *d_beta, double beta_increment, int64_t *d_counts,
                      double *d_scale, mb_per_params *d_params, void *stream);
/* Fixed-shape shard exchange (every (shard, requester) pair always moves `cap` wire rows; d_counts[m] of the
 * rows of block m are real).  All counts are read on the device, so the calls can be replayed in a CUDA graph.
 *  compact:  recv[P][cap][width] -> merged[sum(counts)][width] in member order, rows beyond the sum zeroed up
 *            to merged_rows; metadata split out; is_weight multiplied by *d_scale (0 for padding rows)
 *  scatter:  the write-back of merged row j (block m, position i) -> send[m][i] = (index, version, priority)
 *  mask:     recv[P][cap][24] -> flat index / priority / valid, valid = i < d_counts[self] (rows this shard
 *            served to requester r) and version_table[index] == version (prioritized_buffer_d.py:282-291) */
int mb_wire_compact(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                    const int64_t *d_counts, void *d_merged, int64_t merged_rows, int64_t *d_index,
                    int64_t *d_version, double *d_is_weight, const double *d_scale, void *stream);
int mb_wire_scatter_update(void *d_send, int n_blocks, int64_t cap, const int64_t *d_counts,
                           const int64_t *d_index, const int64_t *d_version, const float *d_priority,
                           int64_t merged_rows, void *stream);
int mb_wire_update_mask(const void *d_recv, int n_blocks, int64_t cap, const int64_t *d_counts, int self_block,
                        const int64_t *d_version_table, int64_t table_size, int64_t *d_index, float *d_priority,
                        uint8_t *d_valid, void *stream);

/* ------------------------------------------------------------------------------------------
 * Shard exchange over peer memory (NVLink / NVSwitch): the collectives of the fixed-shape update as direct
 * stores from this library's kernels into the other GPUs' buffers, instead of NCCL calls.
 * Replaces, on the data path, the per-member RPCs of prioritized_buffer_d.py:221-245 (sample), :184-202 (priority
 * write-back) and the gradient / weight movement of apex.py:141-150.  The buffers are symmetric allocations (same
 * size on every member, peer-mapped once by the host side -- torch.distributed._symmetric_memory in
 * machin_b200/parallel/peer.py); `peers[m]` is member m's buffer as mapped into THIS process (peers[self] = the
 * local one).  Every call takes host arrays of n_blocks device pointers (baked into the launch: capturable).
 *  push_rows:    the cap rows drawn for requester r (block r of d_index / d_is_weight) -> peers[r], block `self`:
 *                [packed row | index | version_table[index] | is_weight] at `width` bytes per row (gather + metadata
 *                + all-to-all in one kernel: the row never exists in a local send buffer)
 *  push_update:  write-back of merged row j (served by shard m at position i) -> peers[m], block `self`, as
 *                (index, version, priority) -- the scatter + all-to-all of the NCCL path
 *  push:         nbytes of d_src -> peers[r] + self * slot_stride for every r (all-gather of shard sums: 8 bytes;
 *                gradient all-reduce, first half: the flat gradient buffer into slot `self` of every member)
 *  sum:          d_out[i] = (slots[0][i] + slots[1][i] + ...) / divisor for i < n, slot m at d_slots + m * slot_floats,
 *                member order -- every member adds the
 *                same numbers in the same order, so all members hold bit-identical averaged gradients
 *  barrier:      every member stores epoch+1 into flag `self` of every member (release, system scope) and waits
 *                until its own n_blocks flags reached it (acquire): what was pushed before the barrier by any
 *                member is visible after it.  d_state[0] = epoch (kept on the device: graph replays advance it),
 *                d_state[1] != 0 after a wait longer than timeout_ns (a member died): checked by the host side. */
#define MB_MAX_PEERS 16
int mb_wire_push_rows(const void *d_ring, int64_t pitch, int64_t capacity, const int64_t *d_index,
                      const double *d_is_weight, const int64_t *d_version_table, int n_blocks, int64_t cap,
                      int self_block, int64_t width, void *const *peers, void *const *peer_flags, unsigned int *d_ticket,
                      uint64_t *d_state, void *stream);
int mb_wire_push_update(int n_blocks, int64_t cap, int self_block, const int64_t *d_counts, const int64_t *d_index,
                        const int64_t *d_version, const float *d_priority, int64_t merged_rows, void *const *peers,
                        void *stream);
int mb_peer_push(const void *d_src, int64_t nbytes, int n_blocks, int self_block, int64_t slot_stride,
                 void *const *peers, void *const *peer_flags, unsigned int *d_ticket, uint64_t *d_state, void *stream);
int mb_peer_sum(const float *d_slots, int n_blocks, int64_t n, int64_t slot_floats, float divisor, float *d_out,
                void *const *peer_flags, int self_block, uint64_t *d_state, void *stream);
/* The barrier folded into its neighbours (peer_flags != NULL): a producer (push_rows, push) publishes the site's next
 * epoch from the LAST CTA of its grid (d_ticket: one zeroed unsigned per site) after fencing its peer stores; a consumer
 * (compact_wait, sum) waits at its top until this member's n_blocks flags reached the epoch in d_state[0].  Saves the
 * separate barrier launch between them.  sums_apportion: all-gather of the shard sums + barrier + mb_wire_apportion in
 * one warp (the first exchange of an update). */
int mb_wire_compact_wait(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                         const int64_t *d_counts, void *d_merged, int64_t merged_rows, int64_t *d_index, int64_t *d_version,
                         double *d_is_weight, const double *d_scale, void *const *peer_flags, int self_block,
                         uint64_t *d_state, void *stream);
/* compact + unpack in one launch: merged row j goes straight into the dense field tensors (`fields` as in mb_rows_gather;
 * row offsets relative to the wire row), padding rows zero-filled, metadata as in mb_wire_compact */
int mb_wire_compact_fields(const void *d_recv, int n_blocks, int64_t cap, int64_t width, int64_t meta_offset,
                           const int64_t *d_counts, const mb_field *fields, int n_fields, int64_t merged_rows,
                           int64_t *d_index, int64_t *d_version, double *d_is_weight, const double *d_scale,
                           void *const *peer_flags, int self_block, uint64_t *d_state, void *stream);
int mb_peer_sums_apportion(void *const *peer_flags, void *const *peer_sums, const double *d_my_sum, int n_blocks,
                           int self_block, uint64_t *d_state, int64_t timeout_ns, int64_t batch, int64_t merged_rows,
                           const int64_t *d_n_stored, double *d_beta, double beta_increment, int64_t *d_counts,
                           double *d_scale, mb_per_params *d_params, void *stream);
int mb_peer_barrier(void *const *peer_flags, int n_blocks, int self_block, uint64_t *d_state,
                    int64_t timeout_ns, void *stream);

/* ------------------------------------------------------------------------------------------
 * TD target / priority math
 * ------------------------------------------------------------------------------------------ */
/* DQNPer.update value block (dqn_per.py:105-139, dqn.py:497-501), float32:
 *   a* = argmax_a q_next[b,:] (first maximum) ; y = reward + (discount * !terminal) * q_tgt_next[b,a*]
 *   action_value = q[b, action[b]] ; abs_error = |action_value - y|
 *   loss_elem = (action_value - y)^2 * is_weight (float64, MSE criterion) ; d_q = dloss/dq for
 *   loss = mean(loss_elem) (nullable).  mode: 0 = "double", 1 = "fixed_target"/"vanilla"
 *   (a* = argmax of q_tgt_next).  Any output pointer may be NULL. */
int mb_td_dqn(const float *d_q, const float *d_q_next, const float *d_q_tgt_next, int64_t batch,
              int n_actions, const int64_t *d_action, const float *d_reward,
              const uint8_t *d_terminal, const double *d_is_weight, float discount, int mode,
              float *d_y, float *d_action_value, float *d_abs_error, double *d_loss_elem,
              float *d_grad_q, void *stream);
/* RAINBOW.update value block (rainbow.py:212-306 + dqn.py:497-501), float32.  Inputs are the
 * soft-maxed value distributions [batch, n_actions, n_atoms] of the online net on s and s' and of the target net
 * on s', d_support = linspace(v_min, v_max, n_atoms) as float32, d_value = the n-step value stored with the
 * transition, discount_n = discount ** reward_future_steps.  Per row: a* = argmax_a sum_k support[k] * next[b,a,k]
 * (first maximum); T_z = clamp(value + (discount_n * !terminal) * support, v_min + 1e-6, v_max - 1e-6);
 * b = (T_z - v_min) / delta_z; target_dist = the two index_add_ passes of :279-280 in element order (a target
 * exactly on an atom loses its mass, as in the reference); loss_row = -sum_k target_dist * log(q[b,action,k] + 1e-6);
 * abs_error = |loss_row| + 1e-6; loss_elem (float64): mean over the batch = the reference's scalar loss, i.e. the
 * mean of the [B, B] product of :303; d_grad_q = d loss / d q_dist.  d_weight_scale: NULL, or one device double
 * = sum(is_weight) / batch^2 supplied by the caller (large batches; otherwise every row's CTA sums d_is_weight).
 * n_atoms <= 256.  Outputs may be NULL. */
int mb_td_rainbow(const float *d_q_dist, const float *d_next_q_dist, const float *d_target_next_q_dist, int64_t batch,
                  int n_actions, int n_atoms, const float *d_support, const int64_t *d_action, const float *d_value,
                  const uint8_t *d_terminal, const double *d_is_weight, const double *d_weight_scale, double v_min,
                  double v_max, double discount_n, float *d_target_dist, float *d_loss_row, float *d_abs_error,
                  double *d_loss_elem, float *d_grad_q, void *stream);
/* DDPGPer target + priority (ddpg_per.py:117-160, ddpg.py:506-510):
 *   y = reward + (discount * !terminal) * next_value ; abs_error = |value - y| (value nullable) */
int mb_td_target(const float *d_next_value, const float *d_reward, const uint8_t *d_terminal,
                 int64_t batch, float discount, const float *d_value, float *d_y,
                 float *d_abs_error, void *stream);
/* soft_update (algorithms/utils.py:8-27): target = target * (1 - rate) + source * rate, over a
 * flat float32 parameter block */
int mb_soft_update(float *d_target, const float *d_source, int64_t n, double rate, void *stream);
/* the same over every parameter tensor of a network in one launch per MB_MAX_SOFT_TENSORS tensors;
 * d_targets / d_sources / sizes are HOST arrays of n_tensors device pointers / element counts */
#define MB_MAX_SOFT_TENSORS 24
int mb_soft_update_multi(float *const *d_targets, const float *const *d_sources, const int64_t *sizes,
                         int n_tensors, double rate, void *stream);

/* ------------------------------------------------------------------------------------------
 * Value-network math for plain Linear+ReLU stacks (float32, fixed summation order).  The reference
 * leaves Q / critic networks to the user's nn.Module (dqn_per.py:105-113); this is the optional fused
 * path for the MLPs its tests and examples use (test_apex.py:25-40).  All matrices row-major.
 * d_workspace (floats) / d_counters (zeroed ints, one per 64x64 output tile) enable split-K; pass NULL to
 * disable it.
 * ------------------------------------------------------------------------------------------ */
/* y[M,N] = x[M,K] * w[N,K]^T + b[N], optional ReLU          (nn.Linear + ReLU) */
int mb_mlp_linear_forward(const float *d_x, const float *d_w, const float *d_b, float *d_y, int M, int N,
                          int K, int relu, float *d_workspace, int64_t workspace_floats, int *d_counters,
                          int n_counters, void *stream);
/* dx[M,K] = dy[M,N] * w[N,K], zeroed where d_act_mask[M,K] <= 0 (ReLU backward; mask nullable) */
int mb_mlp_linear_backward_input(const float *d_dy, const float *d_w, const float *d_act_mask, float *d_dx,
                                 int M, int N, int K, float *d_workspace, int64_t workspace_floats,
                                 int *d_counters, int n_counters, void *stream);
/* dw[N,K] = dy[M,N]^T * x[M,K] ; db[N] = column sums of dy (nullable) */
int mb_mlp_linear_backward_weight(const float *d_dy, const float *d_x, float *d_dw, float *d_db, int M, int N,
                                  int K, float *d_workspace, int64_t workspace_floats, int *d_counters,
                                  int n_counters, void *stream);
/* The whole double-DQN value math of one learner update for Linear+ReLU Q-nets -- forward of Q(s), Q(s'),
 * Q_target(s'), the TD row math of mb_td_dqn, and the backward pass into d_grads -- as ONE cooperative
 * launch (a persistent kernel walks a list of GEMM work items; phases are separated by grid syncs).
 * sizes: n_layers + 1 widths; d_params / d_target_params / d_grads: HOST arrays of 2 * n_layers device
 * pointers (w0, b0, w1, b1, ...).  All device pointers are captured at creation and must stay valid
 * (static batch buffers: CUDA-graph friendly).  Replaces dqn_per.py:105-139 + the user MLP's autograd. */
typedef struct mb_mlp_program mb_mlp_program;
int mb_dqn_mlp_program_create(int device, int n_layers, const int *sizes, int batch, float *const *d_params,
                              float *const *d_target_params, float *const *d_grads, const float *d_state,
                              const float *d_next_state, const int64_t *d_action, const float *d_reward,
                              const uint8_t *d_terminal, const double *d_is_weight, float discount, int mode,
                              float *d_y, float *d_action_value, float *d_abs_error, double *d_loss_elem,
                              mb_mlp_program **out);
/* Same contract and arguments, no grid-wide barrier: the forward passes, the TD row math and the backward-input
 * chain of a batch row depend on that row only, so one CTA per 8 batch rows runs them back to back with the
 * activations in shared memory and the weights streaming from L2; dW / db follow as one launch of independent
 * tile x K-split items.  mb_mlp_program_run issues three launches (weight transposes, row tiles, dW).  Layer
 * outputs wider than 256 are rejected with MB_ERR_INVALID (use mb_dqn_mlp_program_create). */
int mb_dqn_mlp_rows_create(int device, int n_layers, const int *sizes, int batch, float *const *d_params,
                           float *const *d_target_params, float *const *d_grads, const float *d_state,
                           const float *d_next_state, const int64_t *d_action, const float *d_reward,
                           const uint8_t *d_terminal, const double *d_is_weight, float discount, int mode,
                           float *d_y, float *d_action_value, float *d_abs_error, double *d_loss_elem,
                           mb_mlp_program **out);
int mb_mlp_program_run(mb_mlp_program *program, void *stream);
/* The same in three parts, for callers that overlap other work on a second stream: part 0 = refresh of the k-major
 * weight copies (row-tile program only; needs nothing but the weights, so it can run beside the draw and the gather),
 * part 1 = forward passes + TD + backward-input (everything, for the cooperative program), part 2 = weight / bias
 * gradients (row-tile program only; the priority write-back only needs |delta| and can start before it). */
int mb_mlp_program_run_part(mb_mlp_program *program, int part, void *stream);
int mb_mlp_program_destroy(mb_mlp_program *program);
/* torch.optim.Adam step (amsgrad off) over n_tensors parameter tensors in one launch; d_step_count is a
 * device float incremented by the call (capturable in a CUDA graph).  Pointer arrays live on the HOST. */
int mb_adam_step_multi(float *const *d_params, const float *const *d_grads, float *const *d_exp_avg,
                       float *const *d_exp_avg_sq, const int64_t *sizes, int n_tensors, float *d_step_count,
                       double lr, double beta1, double beta2, double eps, double weight_decay, void *stream);
/* The same step followed, element by element, by the soft target update target = target * (1 - rate) + param * rate
 * (utils.py:8-27) on the matching target-network tensors: one pass over the parameters, and the step counter is
 * advanced by the kernel itself (d_ticket: one zero-initialised unsigned of device memory) -- two launches less.
 * At most MB_MAX_SOFT_TENSORS tensors per call. */
int mb_adam_soft_step_multi(float *const *d_params, const float *const *d_grads, float *const *d_exp_avg,
                            float *const *d_exp_avg_sq, float *const *d_target_params, const int64_t *sizes,
                            int n_tensors, float *d_step_count, unsigned *d_ticket, double lr, double beta1,
                            double beta2, double eps, double weight_decay, double update_rate, void *stream);

/* ------------------------------------------------------------------------------------------
 * Trajectory reverse scans
 * ------------------------------------------------------------------------------------------ */
/* IMPALA V-trace (impala.py:322-362), float32.  Steps are flattened episode-major; episode e
 * covers [d_offsets[e], d_offsets[e+1]).  d_offsets == NULL means n_episodes equal episodes of
 * n_steps / n_episodes steps.
 *   sim = exp(log_pi - log_mu) ; c = min(clip_c, sim) ; rho = min(clip_rho, sim)
 *   delta = rho * (reward + discount * (terminal ? 0 : next_value) - value)
 *   vs[i] = value[i] + delta[i] + discount * c[i] * (vs[i+1] - value[i+1]),  vs[last] = 0
 *   vss[i] = vs[i+1], vss[last] = 0 ; adv = reward + discount * vss - value
 * Outputs (each nullable): vs, vss, rho, adv. */
int mb_vtrace(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const float *d_value,
              const float *d_next_value, const float *d_reward, const uint8_t *d_terminal,
              const float *d_log_pi, const float *d_log_mu, float discount, float clip_c,
              float clip_rho, float *d_vs, float *d_vss, float *d_rho, float *d_adv, void *stream);
/* The same with the length of the longest episode given (0 = unknown): lets the launch put fewer episodes on a
 * CTA so that long episodes keep the coalesced, shared-memory-staged path instead of the thread-per-episode one. */
int mb_vtrace_len(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, int64_t max_episode_len,
                  const float *d_value, const float *d_next_value, const float *d_reward,
                  const uint8_t *d_terminal, const float *d_log_pi, const float *d_log_mu, float discount,
                  float clip_c, float clip_rho, float *d_vs, float *d_vss, float *d_rho, float *d_adv, void *stream);
/* A2C/A3C discounted return + GAE (a2c.py:275-313), float64 like the reference's Python floats:
 *   ret[last] = r[last] ; ret[i-1] = ret[i] * discount + r[i-1]
 *   lambda == 1: gae = ret - V(s) ; lambda == 0: gae = r + discount*(1-term)*V(s') - V(s)
 *   else reverse recurrence with V_prev / A_prev (a2c.py:296-313).
 * critic values are float32 (what critic(state).item() widens).  */
int mb_gae(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, const double *d_reward,
           const uint8_t *d_terminal, const float *d_critic_value,
           const float *d_next_critic_value, double discount, double gae_lambda,
           double *d_returns, double *d_gae, void *stream);
int mb_gae_len(const int64_t *d_offsets, int64_t n_episodes, int64_t n_steps, int64_t max_episode_len,
               const double *d_reward, const uint8_t *d_terminal, const float *d_critic_value,
               const float *d_next_critic_value, double discount, double gae_lambda, double *d_returns,
               double *d_gae, void *stream);

int mb_vtrace_host(int device, const int64_t *h_offsets, int64_t n_episodes, int64_t n_steps,
                   const float *h_value, const float *h_next_value, const float *h_reward,
                   const uint8_t *h_terminal, const float *h_log_pi, const float *h_log_mu,
                   float discount, float clip_c, float clip_rho, float *h_vs, float *h_vss,
                   float *h_rho, float *h_adv);
int mb_gae_host(int device, const int64_t *h_offsets, int64_t n_episodes, int64_t n_steps,
                const double *h_reward, const uint8_t *h_terminal, const float *h_critic_value,
                const float *h_next_critic_value, double discount, double gae_lambda,
                double *h_returns, double *h_gae);

#ifdef __cplusplus
}
#endif
#endif /* MACHIN_B200_H */
