/*
 * vivae_b200 -- C ABI of the B200-native (sm_100a) ViVAE hot path.
 *
 * The reference (saeyslab/ViVAE) is pure Python and has no FFI layer; its
 * boundary for this path is three Python call sites.  Each entry point below
 * names the reference interface it replaces (file:line in /root/reference).
 * The binding a reference maintainer would add is a ctypes stub; see
 * INTEGRATION.md.  Plain pointers and sizes only -- no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VVB_E* code;
 *     vvb_last_error() gives the message for the calling thread.
 *   - "*_host" functions take HOST pointers, do their own device allocation,
 *     H2D / D2H copies and synchronise before returning (the drop-in path).
 *   - "*_device" functions take DEVICE pointers on the current CUDA device,
 *     enqueue work on `stream` (a cudaStream_t passed as void*, NULL = legacy
 *     default stream) and do NOT synchronise unless stated.
 *   - matrices are dense row-major; `n` rows, `d` columns.
 */
#ifndef VIVAE_B200_H
#define VIVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VVB_OK 0
#define VVB_EINVAL (-1)   /* bad argument (maps to Python ValueError)            */
#define VVB_ECUDA (-2)    /* CUDA runtime error (maps to Python RuntimeError)    */
#define VVB_ENOMEM (-3)   /* host or device allocation failed                    */
#define VVB_ENOTSUP (-4)  /* valid request outside the implemented envelope      */

/* kNN algorithm selector */
#define VVB_KNN_AUTO 0    /* tensor-core screening + FP32 re-rank + certify/fallback */
#define VVB_KNN_BRUTE 1   /* FP32 FFMA brute force (the exactness fallback, usable on its own) */
#define VVB_KNN_TENSOR 2  /* force the tensor-core path (error if unsupported shape) */

/* smooth() sweep semantics */
#define VVB_SMOOTH_GAUSS_SEIDEL 0 /* the reference's in-place sweep (knn.py:173-184) */
#define VVB_SMOOTH_JACOBI 1       /* every row reads the previous sweep (NOT the reference) */

/* quartet-loss distance functions (losses.py:140-177) */
#define VVB_DIST_EUCLIDEAN 0
#define VVB_DIST_COSINE 1

/* counters filled by the kNN entry points (all optional: pass NULL) */
typedef struct vvb_knn_stats {
  int64_t rows;            /* query rows processed                                  */
  int64_t rows_fallback;   /* rows recomputed by the FP32 brute-force kernel (their
                              certificate failed AND their refine list overflowed)  */
  int32_t algo_used;       /* VVB_KNN_BRUTE or VVB_KNN_TENSOR                       */
  int32_t candidates;      /* m: candidates kept per row by the screening kernel    */
  float ms_prepare;        /* device time, preparation (centre/scale/split, cell plan, operands) */
  float ms_screen;         /* device time, screening launch + join of the half lists */
  float ms_rerank;         /* device time, FP32 re-rank + certify                   */
  float ms_fallback;       /* device time, refine pass + brute-force fallback       */
  int64_t kernel_launches; /* kernels launched by this call                         */
  int64_t tiles_swept;     /* (256 query rows x 128 database rows) tiles the screening
                              kernel actually computed                              */
  int64_t tiles_dense;     /* tiles of the full cross product (what a sweep without
                              the cell bounds would compute)                        */
  int32_t cells;           /* spatial cells of the partition (1 = none)             */
  int32_t rows_refined;    /* rows whose certificate failed and were settled by the
                              fixed-threshold refine sweep on the tensor cores      */
  int32_t products;        /* fp16 products the main sweep executed per key: 3 (hi/lo split)
                              or 1 (single-product first tier of wide rows)         */
  int32_t tiles_pre;       /* of tiles_swept: tiles of the threshold pre-pass that executed ONE product
                              per key (resident variant; saturates at INT32_MAX)       */
} vvb_knn_stats;

int vvb_version(void);
const char *vvb_last_error(void);
/* number of kernels launched by this library in the calling process so far */
int64_t vvb_launch_count(void);
/* the *_host entry points keep their device buffers and streams between calls (grow-only);
 * this frees them */
int vvb_release_cached_memory(void);

/* The two row scans of ensure_valid_knn (/root/reference/src/vivae/knn.py:58-61): is idx[i, 0] == i for
 * every row, and is dist[i, 0] == 0 for every row.  Host arrays, row-major with leading dimensions
 * ld_idx / ld_dist (elements); dist is float32 (dist_elem_size 4) or float64 (8).  No GPU involved. */
int vvb_knn_validate_host(const int64_t *idx_host, int64_t ld_idx, const void *dist_host, int dist_elem_size,
                          int64_t ld_dist, int64_t n, int *self_first_out, int *dist0_zero_out);

/* Page-locked host memory for result arrays.  The *_host entry points recognise page-locked host
 * pointers (these, or any cudaHostAlloc / cudaHostRegister memory the caller owns) and DMA to / from
 * them directly; pageable pointers are staged through bounce buffers.  The Python layer returns its
 * result arrays (knn.py:137-141, 185) in recycled blocks from here. */
int vvb_host_alloc(size_t bytes, void **ptr_out);
int vvb_host_free(void *ptr);

/* ------------------------------------------------------------------------
 * make_knn  -- replaces the cache-miss branch of
 *   /root/reference/src/vivae/knn.py:133-141
 *   (pynndescent.NNDescent(x, n_neighbors=k+1).query(x, k=k+1), int64 cast,
 *    correct_knn_for_duplicates), made EXACT:
 *   for every query row i, all rows j != i ranked by the FP32 key (d2(i,j), j),
 *   d2 = left-to-right fmaf chain over coordinates of (x_ic - x_jc)^2;
 *   column 0 is i itself with distance 0; dist = sqrtf(d2).
 *
 *   x        (n, d) float32, row-major           -- the database AND the query set
 *   q0, nq   query rows [q0, q0+nq) (row sharding; q0=0,nq=n for everything)
 *   idx_out  (nq, k+1) int64   -- global row indices, like knn.py:137
 *   dist_out (nq, k+1) float32 -- knn.py:135-136 leaves distances float32
 * ------------------------------------------------------------------------ */
int vvb_make_knn_host(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq,
                      int algo, int64_t *idx_out_host, float *dist_out_host, vvb_knn_stats *stats);

/* bytes of device workspace vvb_make_knn_device needs for these sizes */
int vvb_make_knn_workspace_bytes(int64_t n, int d, int k, int64_t nq, int algo, size_t *bytes);

/* same contract with device-resident buffers.  The tensor path reads the number of rows its
 * certificate rejected back to the host once per chunk of up to 1,048,576 rows (a stream synchronisation), so
 * the call is not capturable in a CUDA graph; on return all work is enqueued (and, with `stats`
 * non-NULL, finished). */
int vvb_make_knn_device(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq,
                        int algo, int64_t *idx_out_dev, float *dist_out_dev, void *workspace_dev,
                        size_t workspace_bytes, void *stream, vvb_knn_stats *stats);

/* ------------------------------------------------------------------------
 * Multi-GPU preparation of the tensor path (same reference call site, knn.py:135-141; the reference is
 * single-device).  With the matrix replicated and the QUERY rows sharded over `parts` GPUs, every GPU would
 * repeat the whole preparation (pivot assignment and cell extents are O(n P) each); instead part `part` runs
 * vvb_knn_plan_phase_device for phase = 0, 1, 2, 3 on its own workspace, and after each phase the parts combine
 * the tables it reports in xch[0 .. *n_xch) -- device pointers INTO the workspace -- with the collective `op`
 * names (NCCL in vivae_b200.device.knn_device; any all-gather / all-reduce will do):
 *   VVB_XCH_GATHER_ROWS  element i belongs to the part whose row shard holds row i (shards of ceil(n/parts) rows)
 *   VVB_XCH_SUM          element-wise sum over the parts
 *   VVB_XCH_MAX_ORD      element-wise max over the parts of uint32 values compared as UNSIGNED
 * Then vvb_make_knn_planned_device (arguments of vvb_make_knn_device, same workspace, same n/d/k/nq) builds the
 * graph rows [q0, q0+nq) without preparing again.  parts = 1 needs no exchange and equals vvb_make_knn_device.
 * Returns VVB_ENOTSUP when (n, d, k) is outside the tensor path (use vvb_make_knn_device).
 * ------------------------------------------------------------------------ */
#define VVB_XCH_I32 0
#define VVB_XCH_U32 1
#define VVB_XCH_F64 2
#define VVB_XCH_GATHER_ROWS 0
#define VVB_XCH_SUM 1
#define VVB_XCH_MAX_ORD 2
typedef struct vvb_exchange {
  void *ptr;      /* device pointer into the workspace */
  size_t count;   /* elements */
  int32_t dtype;  /* VVB_XCH_I32 / U32 / F64 */
  int32_t op;     /* VVB_XCH_GATHER_ROWS / SUM / MAX_ORD */
} vvb_exchange;

int vvb_knn_plan_phase_device(const float *x_dev, int64_t n, int d, int k, int64_t nq, void *workspace_dev,
                              size_t workspace_bytes, int phase, int part, int parts, void *stream,
                              vvb_exchange *xch /* [2] */, int *n_xch);

int vvb_make_knn_planned_device(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq,
                                int64_t *idx_out_dev, float *dist_out_dev, void *workspace_dev,
                                size_t workspace_bytes, void *stream, vvb_knn_stats *stats);

/* Diagnostics (no reference counterpart): run only the preparation and the tensor-core
 * screening of the VVB_KNN_TENSOR path and return the raw per-row candidate lists:
 * cand_idx/cand_key (nq, C) with C = vvb_knn_screen_debug_capacity(), cand_cnt (nq) valid
 * entries per row (sorted by screening key), cand_thr (nq) the final admission threshold,
 * xnorm (nq) |x^|^2 of the scaled/centred query, scale_ymax[2] = {scale, ymax}, mu the
 * column means (64 * ceil((d + 3) / 64) floats, zero beyond d).  Tests use it to bound the
 * screening error against FP64. */
int vvb_knn_screen_debug_capacity(void);
int vvb_knn_screen_debug_host(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq,
                              uint32_t *cand_idx_out, float *cand_key_out, int32_t *cand_cnt_out,
                              float *cand_thr_out, float *xnorm_out, float *scale_ymax_out, float *mu_out);

/* ------------------------------------------------------------------------
 * Database-sharded ("ring") graph build: running top-k merge.
 *   Same reference call site as make_knn (knn.py:135-141), for a matrix whose rows are spread
 *   over several GPUs.  A rank keeps, for each of its nq query rows, a running list
 *   run_idx (nq, k+1) int64 GLOBAL row indices and run_d2 (nq, k+1) float32 SQUARED distances
 *   (column 0 = the row itself, unfilled entries -1 / +inf).  For every database shard that
 *   passes by it calls vvb_make_knn_device on the stacked matrix xc = [shard ; own rows]
 *   (stacked in global row order) and then vvb_knn_merge_device, which
 *     - keeps the entries of new_idx (nq, kn+1; local row indices into xc, column 0 = self)
 *       whose index is in [keep_lo, keep_hi) (the shard's rows; the rank's own rows are ranked
 *       once, in its first step),
 *     - gives them global indices (local + global_off) and the contract's squared distance
 *       (recomputed fmaf chain),
 *     - merges them into the running lists by the key (d2, global index) and keeps the k best.
 *   vvb_knn_merge_finish_device turns squared distances into distances (correctly rounded sqrt).
 *   The result is bit-identical to vvb_make_knn_device over the whole matrix.
 * ------------------------------------------------------------------------ */
int vvb_knn_merge_init_device(int64_t *run_idx_dev, float *run_d2_dev, int64_t nq, int k, int64_t self0,
                              void *stream);

int vvb_knn_merge_device(const float *xc_dev, int64_t nc, int d, int64_t q0, int64_t nq,
                         const int64_t *new_idx_dev, int kn, int64_t keep_lo, int64_t keep_hi,
                         int64_t global_off, int64_t *run_idx_dev, float *run_d2_dev, int k, void *stream);

int vvb_knn_merge_finish_device(const float *run_d2_dev, int64_t count, float *dist_out_dev, void *stream);

/* ------------------------------------------------------------------------
 * smooth -- replaces the sweep loop of /root/reference/src/vivae/knn.py:173-185
 *   (argument validation and ensure_valid_knn, knn.py:165-171, stay in Python).
 *   Row i moves by `coef` toward the mean of rows idx[i, 0:k]; in
 *   VVB_SMOOTH_GAUSS_SEIDEL mode rows j < i are read already-updated, exactly
 *   as the reference's aliased in-place loop does.  Arithmetic per element:
 *   sequential sum in neighbour order, true divide by k, then
 *   r + (target - r) * coef as three separately rounded operations.
 *
 *   elem_size 4 = float32, 8 = float64 (dtype follows x, like the reference)
 *   idx (n, ld_idx) int64, ld_idx >= k ; out (n, d) same dtype, must not alias x
 * ------------------------------------------------------------------------ */
int vvb_smooth_host(const void *x_host, int elem_size, int64_t n, int d, const int64_t *idx_host,
                    int64_t ld_idx, int k, double coef, int n_iter, int mode, void *out_host);

int vvb_smooth_workspace_bytes(int elem_size, int64_t n, int d, int n_iter, size_t *bytes);

int vvb_smooth_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_dev,
                      int64_t ld_idx, int k, double coef, int n_iter, int mode, void *out_dev,
                      void *workspace_dev, size_t workspace_bytes, void *stream);

/* Columns [col0, col0+dcols) of the same computation (all n_iter sweeps, either mode).  The smoother never mixes
 * columns -- element (i, c) of a sweep depends on column c of row i's neighbours only (knn.py:176-183 is a row-wise
 * mean) -- so a column slice is bit-identical to the same columns of the full call, Gauss-Seidel order included.
 * This is how several GPUs share the reference's sweep, which does NOT split by rows: every GPU takes a slice and
 * the slices are all-gathered (vivae_b200.sharded.smooth_cols_sharded).  x_dev and out_dev are the FULL (n, d)
 * matrices; only the slice of out_dev is written.  Workspace as vvb_smooth_workspace_bytes(elem_size, n, d, n_iter). */
int vvb_smooth_cols_device(const void *x_dev, int elem_size, int64_t n, int d, int col0, int dcols,
                           const int64_t *idx_dev, int64_t ld_idx, int k, double coef, int n_iter, int mode,
                           void *out_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * The reference's (Gauss-Seidel) sweep over SEVERAL GPUs in one pass -- same call site, knn.py:173-185.
 *   Rows are interleaved over `world` GPUs (row i belongs to rank i mod world); every rank keeps a complete copy of
 *   the sweep's output in memory its peers can write (vvb_peer_alloc / vvb_peer_open: CUDA IPC over NVLink), as one
 *   64-bit word {value bits, sweep epoch} per float32 element (two per float64 element).  vvb_smooth_peer_device runs
 *   ONE sweep on rank `rank`: it computes the rank's rows from its local copies and stores each finished element, value
 *   and epoch in one 8-byte store, into every rank's copy; a row that needs a neighbour still in flight re-reads the
 *   word until the epoch matches.  No flags, no fences.  ll_ptrs: HOST array of `world` device pointers -- every rank's
 *   copy (n * d * 8 bytes for float32, 16 for float64) as mapped into this process; entry [rank] is the local one.
 *   Contract: the buffers are zero after vvb_peer_alloc; `epoch` >= 1 grows by one per sweep; all ranks call it for
 *   the same sweep, with a barrier over the ranks before (nobody still reads the previous sweep) and after (all pushes
 *   landed); then vvb_smooth_peer_unpack_device turns the local copy into plain values.
 *   x_dev is this rank's copy of the sweep's input.  Bit-identical to vvb_smooth_device on one GPU.
 *   k <= 128, d <= 128 (VVB_ENOTSUP otherwise).  workspace: vvb_smooth_peer_workspace_bytes(); its first int is set
 *   non-zero by an out-of-range neighbour index (1) or a word that never arrived (2).
 * ------------------------------------------------------------------------ */
int vvb_peer_alloc(size_t bytes, void **ptr_out, unsigned char *handle_out /* 64 bytes */);
int vvb_peer_open(const unsigned char *handle /* 64 bytes, from another process */, void **ptr_out);
int vvb_peer_close(void *ptr);
int vvb_peer_free(void *ptr);
size_t vvb_smooth_peer_workspace_bytes(void);
int vvb_smooth_peer_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_dev,
                           int64_t ld_idx, int k, double coef, int epoch, int rank, int world,
                           void *const *ll_ptrs, void *workspace_dev, void *stream);
int vvb_smooth_peer_unpack_device(const void *ll_local_dev, int elem_size, int64_t count, int epoch, void *out_dev,
                                  void *workspace_dev, void *stream);

/* Rows [row_lo, row_hi) of ONE sweep in VVB_SMOOTH_JACOBI mode (row ranges of a Jacobi sweep are independent:
 * several GPUs each take a range and exchange the new rows between sweeps).  x (n, d) is the whole matrix;
 * idx_rows (row_hi-row_lo, ld_idx) and out_rows (row_hi-row_lo, d) hold only the range.  Workspace as
 * vvb_smooth_workspace_bytes(elem_size, n, d, 1); its first int is set non-zero by an out-of-range index.
 * VVB_SMOOTH_GAUSS_SEIDEL is refused (VVB_ENOTSUP): a row of the reference's sweep depends on the rows below. */
int vvb_smooth_rows_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_rows_dev,
                           int64_t ld_idx, int k, double coef, int mode, int64_t row_lo, int64_t row_hi,
                           void *out_rows_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * quartet ("stochastic MDS") loss -- replaces MDSLoss.__call__,
 *   /root/reference/src/vivae/losses.py:266-349, and its autograd backward.
 *
 *   x (B, D) float32, z (B, L) float32 on the device.
 *   perms: n_sampling device pointers (array on the HOST) to int64 permutations
 *          of range(B); a NULL entry means identity (losses.py:328-331).
 *   loss_out : device float32 scalar  = mean over passes of (B/nq) * mean_q cost_q
 *   grad_unit: device (B, L) float32 or NULL -- d(loss)/dz for grad_output = 1
 *              (rows 4*nq..B-1 get 0).
 *   workspace: at least vvb_quartet_workspace_bytes(B) bytes.
 *   B < 4 gives NaN like the reference (mean over zero quartets).
 * ------------------------------------------------------------------------ */
size_t vvb_quartet_workspace_bytes(int64_t B);

int vvb_quartet_loss_fwd_device(const float *x_dev, const float *z_dev, int64_t B, int D, int L,
                                const int64_t *const *perms, int n_sampling, int metric_hd,
                                int metric_ld, float *loss_out_dev, float *grad_unit_dev,
                                void *workspace_dev, void *stream);

/* grad_z = grad_unit * (*grad_out_dev)   (grad_out is the upstream 0-dim gradient) */
int vvb_quartet_loss_bwd_device(const float *grad_unit_dev, const float *grad_out_dev, int64_t count,
                                float *grad_z_dev, void *stream);

/* ------------------------------------------------------------------------
 * Fused training step -- replaces, for the default configuration, what one batch of
 *   /root/reference/src/vivae/model.py:130-177 (`train_epoch`) executes between `optimizer.zero_grad()` and
 *   `optimizer.step()`: `Autoencoder.forward` (network.py:219-296: encode, reparameterise 113-130, KLDivLoss
 *   losses.py:105-131, decode, nn.MSELoss, MDSLoss losses.py:266-349 with euclidean distances and n_sampling = 1) and
 *   the autograd backward of `recon + kldiv + mds`.  ONE kernel (+ one deterministic reduction) instead of ~140.
 *
 *   Layers, in this order: n_enc encoder Linear layers (GELU, exact erf, after each), the mu head, the logvar head,
 *   n_dec decoder Linear layers (GELU after all but the last).  w_ptrs / b_ptrs: HOST arrays of n_enc + 2 + n_dec
 *   DEVICE pointers to the layers' weight (out, in) row-major and bias (out) tensors -- the network's own parameter
 *   tensors, read in place; in_dims / out_dims: HOST arrays of the layers' shapes.
 *   x (B, D) the batch, eps (B, L) the reparameterisation noise (drawn by the caller: torch.randn_like, so the RNG
 *   stream is the reference's); B a multiple of 4.
 *   Outputs: grad_flat (n_params): d(loss)/d(parameters), layer by layer, weight then bias;
 *            loss_out[3] = {lam_recon * MSE, lam_kldiv * KL, lam_mds * MDS}; term_sums (6 floats, may be NULL): the
 *            same three values ADDED to slots 0, 1, 4 (the epoch's running sums, model.py:138-143).
 *   Scratch: grad_partials (n_cta * n_params floats), loss_partials (4 * n_cta floats).
 *   Size query: call with grad_partials_dev == NULL; *smem_bytes, *n_params, *n_cta are filled either way.
 *   VVB_ENOTSUP when the activations of 32 rows do not fit shared memory or the layer count exceeds 16.
 * ------------------------------------------------------------------------ */
int vvb_vae_step_device(const float *const *w_ptrs, const float *const *b_ptrs, const int *in_dims, const int *out_dims,
                        int n_enc, int n_dec, const float *x_dev, const float *eps_dev, int B, int D, int L,
                        float lam_recon, float lam_kldiv, float lam_mds, float *grad_partials_dev,
                        float *loss_partials_dev, float *grad_flat_dev, float *term_sums_dev, float *loss_out_dev,
                        size_t *smem_bytes, int *n_params, int *n_cta, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VIVAE_B200_H */
