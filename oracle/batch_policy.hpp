// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU statement of the *batch* update policy that the CUDA library implements
// (DESIGN.md §4).  The reference has no batched update ("Batch" exists
// only as bulk build or a loop of setindex!, matrix.jl:119-121), so the GPU layout
// after a batch cannot equal the reference's one-op-at-a-time layout in general.
// This file defines, on the CPU and with the oracle's literal pack!/spread!
// (moves.jl:94-172), the layout the GPU must produce bit-for-bit:
//   * ops are de-duplicated last-writer-wins and sorted by (partition, key);
//   * hits are overwritten / blanked in place (writes.jl:16-19, 65-68);
//   * every new key is assigned to the leaf of its predecessor cell;
//   * each touched leaf walks leaf->root exactly like _look_for_rebalance!
//     (pma.jl:105-141) on the post-batch counts and marks the first window whose
//     density is inside [p_0+p_d*h, t_0+t_d*h]; nested marks collapse to the
//     outermost one; each marked window is re-laid with pack!+spread!;
//   * a failing root doubles / halves the capacity (pma.jl:143-161) until the
//     root density is inside its thresholds, then the whole array is re-spread.
//   * a batch of exactly ONE op (on an existing partition, or on a vector) is not
//     re-laid at all: it IS the reference's setindex! — shift to the next gap
//     (writes.jl:26-43, moves.jl:7-85), one leaf->root walk, at most one
//     _extend!/_shrink! — so single writes are layout-bit-exact with the reference.
// Logical contents after a batch are identical to applying the ops one by one
// with the reference (tests check this against the sequential oracle).
#pragma once
#include "dsa_oracle.hpp"

namespace orc {
namespace policy {

enum OpKind : int { OP_SET = 0, OP_DEL = 1, OP_SEM = 2 };
struct Op {
    int64_t pid;    // partition id (1-based, new numbering); 0 for a plain PMA
    int64_t key;
    double val;
    int kind;
    int64_t pos;    // 1-based position of the exact hit or of the predecessor cell (0 = none)
    bool hit;
};

// integer count bounds per level, from the same Float64 expressions as pma.jl:120-123
void level_bounds(int64_t segment_capacity, int64_t height, double t_d, double p_d,
                  std::vector<int64_t>& mn, std::vector<int64_t>& mx);

// apply located ops + purge ranges [a,b] (1-based, inclusive); rebalances per the policy above
void apply_located(Pma& p, Semaphores* sem, std::vector<Op>& ops,
                   const std::vector<std::pair<int64_t, int64_t>>& purge_ranges);

void pma_set_batch(Pma& p, const int64_t* keys, const double* vals, int64_t n);
void mpcsc_set_batch(Mpcsc& m, const int64_t* inkeys, const int64_t* partkeys, const double* vals, int64_t n);
// deletecolumn!/deleterow! for a list of partition keys: purge in `primary`, delete the entries from `twin`
void matrix_delete_partitions(Mpcsc& primary, Mpcsc& twin, const int64_t* ids, int64_t n);

}  // namespace policy
}  // namespace orc
