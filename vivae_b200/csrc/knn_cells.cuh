// Spatial cells for the exact kNN screening sweep (see knn_cells.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vvb {

constexpr int CELL_MAX = 1024;   // most cells a plan may have
constexpr int CELL_DIMS = 64;    // geometry uses the first min(d, 64) coordinates
constexpr int CELL_GROUP = 128;  // cells are padded to whole database tiles

struct CellPlan {
  int P;                 // number of cells (1 = no partition)
  int dsub;              // coordinates used by the geometry
  int64_t n;             // database rows
  int64_t n_perm_cap;    // rows of the permuted (cell-sorted, padded) order, upper bound, multiple of 128
  // device buffers
  int *cell_of;          // (n) cell of every row
  int *cnt;              // (P) rows per cell
  int *off;              // (P + 1) first permuted row of every cell (multiples of 128); off[P] = rows in use
  int *cursor;           // (P) scatter cursors
  int *perm;             // (n_perm_cap) permuted position -> original row, -1 = padding
  double *sums;          // (P, 64) per-cell coordinate sums of the centred rows
  double *sqs;           // (P, 64) per-cell sums of squares
  float *cen;            // (P, 64) cell centres (centred coordinates)
  float *var;            // (64) pooled within-cell variance per coordinate
  float *cw;             // (P, 64) centres / variance
  float *inv_norm;       // (P, P) 1 / |(c_b - c_a) / var|
  float *gap;            // (P, P) projection of c_b - c_a on the unit discriminant direction a -> b
  float *cdist;          // (P, P) |c_b - c_a|
  unsigned int *ext;     // (P, P) ordered-int image of max over rows x of cell a of u_ab . (x - c_a)
  unsigned int *rad2;    // (P) ordered-int image of the largest squared distance of a member to its centre
  float *lb;             // (P, P) safe lower bound of the SCALED distance between any row of a and any row of b
  // per query chunk
  int *qcnt, *qoff, *qcursor;  // (P), (P + 1), (P)
};

int cells_choose_count(int64_t n);
int64_t cells_perm_capacity(int64_t n, int P);
// carve the plan's buffers out of `ws` (nullptr: size query); returns bytes used
size_t cells_carve(CellPlan *plan, void *ws, int64_t n, int d, int P);
// partition the rows, sort them by cell (perm) and compute the pairwise cell bounds.
// mu: column means (>= 64 entries, zero beyond d); scale_dev: device pointer to {power-of-two scale, ymax}
// dims: CELL_DIMS column indices (cells_select_dims)
int cells_build(const CellPlan &plan, const float *x_dev, int d, const float *mu_dev, const int *dims,
                const float *scale_dev, cudaStream_t st);
// one phase (0..3) of the same build for part `part` of `parts` (multi-GPU: see knn_cells.cu)
int cells_build_phase(const CellPlan &plan, const float *x_dev, int d, const float *mu_dev, const int *dims,
                      const float *scale_dev, int phase, int part, int parts, int64_t row_lo, int64_t row_hi,
                      cudaStream_t st);
// the min(d, 64) columns of largest variance; `partial`: nblocks * ceil(d / 64) * 64 doubles of scratch
int cells_select_dims(const float *x_dev, int64_t n, int d, const float *mu_dev, double *partial, int nblocks,
                      int *dims, cudaStream_t st);
// cell-sorted list of the query rows [q0, q0 + nq): qlist[0 .. nq) original row ids, qlist[nq .. nq_pad) = -1
int cells_query_list(const CellPlan &plan, int64_t q0, int64_t nq, int64_t nq_pad, int *qlist, const int64_t *row_list,
                     cudaStream_t st);

// launch order of the query blocks, longest estimated sweep first (knn_cells.cu)
size_t cells_cta_order_scratch_ints(int64_t n_blocks);
int cells_cta_order(const CellPlan &plan, const int *qlist, int64_t nq, int rows_per_block, int n_blocks, int *scratch,
                    int *order, cudaStream_t st);

}  // namespace vvb
