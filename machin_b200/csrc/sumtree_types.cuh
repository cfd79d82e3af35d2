// Device-side view of a sum tree (shared by sumtree.cu and sumtree_update.cuh).
#pragma once
#include <stdint.h>

struct TreeState {
    double max_leaf;   // running maximum leaf weight (WeightTree.max_leaf)
    double isw_max;    // scratch: max importance weight of the draw in flight
    double batch_max;  // scratch: max weight of the update batch in flight (-inf between launches)
    int err;           // deferred device-side validation error (mb_status), sticky until checked
    int err_gen;       // launch generation that raised it (so later launches are not skipped)
    int max_is_f32;    // max_leaf came from a float32 batch (the reference keeps it as np.float32 then,
                       // which changes the dtype of normalize(max_leaf) for default-priority stores)
    int ticket;        // tree_update_part: CTAs that have finished their subtree (the last one lifts the top; 0 between launches)
};

struct TreeDev {
    double *W;
    int *owner;        // per-leaf "last writer" stamp used to resolve duplicate indexes
    TreeState *st;
    int64_t size;      // number of real leaves
    int64_t total;     // 2^depth
    int depth;
};
