/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the ViVAE hot path.
 *
 * Nothing in the product package (vivae_b200/) links, loads or calls this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status
 *   smooth   : PINNED  -- checked bit-for-bit against the reference function
 *              (/root/reference/src/vivae/knn.py:147-185) by oracle/make_golden.py;
 *              fixtures in tests/golden/smooth_*.npz.
 *   quartet  : PINNED  -- loss and d(loss)/dz checked against the reference
 *              MDSLoss + autograd (/root/reference/src/vivae/losses.py:266-349);
 *              fixtures in tests/golden/quartet_*.npz.
 *   kNN      : PARITY UNPINNED at the reference boundary.  The reference calls
 *              pynndescent==0.5.11 (knn.py:135-136; pin at pyproject.toml:29), an
 *              approximate randomised third-party library that is not vendored
 *              and not installable here, and the reference has no test or golden
 *              vector for it.  The contract restated here is BASELINE.json's:
 *              exact (k+1)-NN under the FP32 key (d2, j), d2 = left-to-right
 *              fmaf chain of squared coordinate differences, self in column 0
 *              with distance 0, dist = sqrtf(d2); output layout follows
 *              knn.py:135-141 (idx int64 (n,k+1), dist float32 (n,k+1)).
 *              Cross-checked against scipy cKDTree / sklearn brute force in
 *              tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp).  Floating
 * point contraction is OFF: every fused multiply-add below is an explicit fmaf.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(VV_NO_CLONES)
#define VV_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define VV_CLONES
#endif

#define VV_BLK 16 /* database rows per transposed block (one AVX-512 vector of float) */

int vv_oracle_version(void) { return 1; }

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use all host cores */
void vv_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int vv_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* exact kNN                                                                  */
/* ------------------------------------------------------------------------- */

/* key order: (d2, j) ascending.  Returns 1 if (a,ja) < (b,jb). */
static inline int key_less(float a, int64_t ja, float b, int64_t jb) {
  return (a < b) || (a == b && ja < jb);
}

/* max-heap on (d2, j): root = current worst of the kept k */
static void heap_sift_down(float *hd, int64_t *hj, int k, int p) {
  for (;;) {
    int l = 2 * p + 1, r = l + 1, m = p;
    if (l < k && key_less(hd[m], hj[m], hd[l], hj[l])) m = l;
    if (r < k && key_less(hd[m], hj[m], hd[r], hj[r])) m = r;
    if (m == p) return;
    float td = hd[p]; hd[p] = hd[m]; hd[m] = td;
    int64_t tj = hj[p]; hj[p] = hj[m]; hj[m] = tj;
    p = m;
  }
}

/* squared distances from one query row to one transposed block of VV_BLK
 * database rows.  Per database row the arithmetic is exactly
 *     acc = 0; for c in 0..d-1: t = q[c] - y[c]; acc = fmaf(t, t, acc)
 * (SURVEY.md section 8(c)); the loop over `l` only vectorises ACROSS rows. */
VV_CLONES
static void block_d2(const float *q, const float *yt, int d, float *acc) {
  for (int l = 0; l < VV_BLK; ++l) acc[l] = 0.0f;
  for (int c = 0; c < d; ++c) {
    const float qc = q[c];
    const float *yc = yt + (size_t)c * VV_BLK;
    for (int l = 0; l < VV_BLK; ++l) {
      float t = qc - yc[l];
      acc[l] = fmaf(t, t, acc[l]);
    }
  }
}

/*
 * Exact (k+1)-NN of query rows [q0, q1) of x against all n rows of x.
 *   x    : (n, d) row-major float32
 *   idx  : (q1-q0, k+1) int64   column 0 = the row itself
 *   dist : (q1-q0, k+1) float32 column 0 = 0.0, then ascending sqrtf(d2)
 * Returns 0 on success, <0 on bad arguments / allocation failure.
 */
int vv_oracle_knn_f32(const float *x, int64_t n, int d, int k, int64_t q0, int64_t q1,
                      int64_t *idx, float *dist) {
  if (!x || !idx || !dist || n < 2 || d < 1 || k < 1 || k > n - 1 || q0 < 0 || q1 > n || q0 > q1)
    return -1;
  const int64_t nblk = (n + VV_BLK - 1) / VV_BLK;
  float *xt = (float *)malloc((size_t)nblk * d * VV_BLK * sizeof(float));
  if (!xt) return -2;
  /* transposed copy: xt[b][c][l] = x[b*VV_BLK + l][c]; padding rows replicate row n-1
   * (their results are discarded by the j < n test below) */
  for (int64_t b = 0; b < nblk; ++b)
    for (int c = 0; c < d; ++c)
      for (int l = 0; l < VV_BLK; ++l) {
        int64_t j = b * VV_BLK + l;
        if (j >= n) j = n - 1;
        xt[((size_t)b * d + c) * VV_BLK + l] = x[(size_t)j * d + c];
      }
  int err = 0;
#pragma omp parallel
  {
    float *hd = (float *)malloc((size_t)k * sizeof(float));
    int64_t *hj = (int64_t *)malloc((size_t)k * sizeof(int64_t));
    if (!hd || !hj) {
#pragma omp atomic write
      err = -2;
    }
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = q0; i < q1; ++i) {
      if (err) continue;
      const float *q = x + (size_t)i * d;
      int cnt = 0;
      float acc[VV_BLK];
      for (int64_t b = 0; b < nblk; ++b) {
        block_d2(q, xt + (size_t)b * d * VV_BLK, d, acc);
        for (int l = 0; l < VV_BLK; ++l) {
          const int64_t j = b * VV_BLK + l;
          if (j >= n || j == i) continue; /* self is forced into column 0 below */
          const float v = acc[l];
          if (cnt < k) {
            hd[cnt] = v; hj[cnt] = j; ++cnt;
            if (cnt == k)
              for (int p = k / 2 - 1; p >= 0; --p) heap_sift_down(hd, hj, k, p);
          } else if (key_less(v, j, hd[0], hj[0])) {
            hd[0] = v; hj[0] = j;
            heap_sift_down(hd, hj, k, 0);
          }
        }
      }
      /* heap-sort ascending into the output row */
      int64_t *oi = idx + (size_t)(i - q0) * (k + 1);
      float *od = dist + (size_t)(i - q0) * (k + 1);
      oi[0] = i; od[0] = 0.0f;
      for (int m = k; m > 0; --m) {
        oi[m] = hj[0]; od[m] = sqrtf(hd[0]);
        hd[0] = hd[m - 1]; hj[0] = hj[m - 1];
        heap_sift_down(hd, hj, m - 1, 0);
      }
    }
    free(hd); free(hj);
  }
  free(xt);
  return err;
}

/* ------------------------------------------------------------------------- */
/* smooth  (reference: /root/reference/src/vivae/knn.py:147-185)              */
/* ------------------------------------------------------------------------- */
/*
 * knn.py:173-175 : y = deepcopy(x); res = y            -> ONE working array
 * knn.py:177-183 : for i in 0..n-1:
 *                    target = y[idx[i, 0:k]].mean(0)    -> sequential sum over the k
 *                                                          gathered rows in column order,
 *                                                          then a true divide by k
 *                    res[i] = res[i] + (target - res[i]) * coef   (3 rounded ops)
 * Because res aliases y, row i sees rows j < i already updated in this sweep
 * (Gauss-Seidel).  mode 1 = Jacobi (every row reads the previous sweep) -- the
 * documented non-parity variant of the product, provided here so it can be tested.
 */
#define VV_DEFINE_SMOOTH(NAME, T)                                                              \
  int NAME(const T *x, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k, double coef, \
           int n_iter, int mode, T *out) {                                                     \
    if (!x || !idx || !out || n < 1 || d < 1 || k < 1 || ld_idx < k || n_iter < 1) return -1;  \
    const T c = (T)coef, kk = (T)k;                                                            \
    T *tmp = (T *)malloc((size_t)d * sizeof(T));                                               \
    T *prev = NULL;                                                                            \
    if (!tmp) return -2;                                                                       \
    memcpy(out, x, (size_t)n * d * sizeof(T));                                                 \
    if (mode == 1) {                                                                           \
      prev = (T *)malloc((size_t)n * d * sizeof(T));                                           \
      if (!prev) { free(tmp); return -2; }                                                     \
    }                                                                                          \
    for (int it = 0; it < n_iter; ++it) {                                                      \
      const T *src = out;                                                                      \
      if (mode == 1) { memcpy(prev, out, (size_t)n * d * sizeof(T)); src = prev; }             \
      for (int64_t i = 0; i < n; ++i) {                                                        \
        const int64_t *nb = idx + (size_t)i * ld_idx;                                          \
        for (int cc = 0; cc < d; ++cc) tmp[cc] = src[(size_t)nb[0] * d + cc];                  \
        for (int jj = 1; jj < k; ++jj) {                                                       \
          const T *row = src + (size_t)nb[jj] * d;                                             \
          for (int cc = 0; cc < d; ++cc) tmp[cc] = tmp[cc] + row[cc];                          \
        }                                                                                      \
        T *ri = out + (size_t)i * d;                                                           \
        const T *ro = src + (size_t)i * d;                                                     \
        for (int cc = 0; cc < d; ++cc) {                                                       \
          T target = tmp[cc] / kk;                                                             \
          T vec = target - ro[cc];                                                             \
          T prod = vec * c;                                                                    \
          ri[cc] = ro[cc] + prod;                                                              \
        }                                                                                      \
      }                                                                                        \
    }                                                                                          \
    free(tmp); free(prev);                                                                     \
    return 0;                                                                                  \
  }

VV_DEFINE_SMOOTH(vv_oracle_smooth_f32, float)
VV_DEFINE_SMOOTH(vv_oracle_smooth_f64, double)

/* ------------------------------------------------------------------------- */
/* quartet ("stochastic MDS") loss                                            */
/* reference: /root/reference/src/vivae/losses.py:266-349                     */
/* ------------------------------------------------------------------------- */
/*
 * One sampling pass (losses.py:326-343).  Rows of quartet q are perm[a*nq + q],
 * a = 0..3 (losses.py:326,332-336; perm == NULL means identity, losses.py:331).
 * Pairs in the order of losses.py:234,263: (0,1) (1,2) (2,3) (0,2) (0,3) (1,3).
 * metric 0: euclidean  sqrt(max(sum (a-b)^2, 1e-9))            losses.py:154-159
 * metric 1: cosine     1 - sum (a/max(|a|,1e-8)) (b/max(|b|,1e-8))   losses.py:176
 * Pass loss = (B/nq) * mean_q sum_pairs (dx/X - dz/S)^2         losses.py:263,339-343
 * Gradient w.r.t. z in closed form; x has no gradient.
 * grad_z (B, L) is ACCUMULATED into (caller zeroes it); returns the pass loss.
 * All arithmetic in double: this is the tolerance oracle (1e-4 relative), the
 * FP32 behaviour of the reference is captured by the golden fixtures.
 */
static const int VV_PA[6] = {0, 1, 2, 0, 0, 1};
static const int VV_PB[6] = {1, 2, 3, 2, 3, 3};

static double pair_dist(const double *a, const double *b, int dim, int metric) {
  if (metric == 0) {
    double s = 0.0;
    for (int c = 0; c < dim; ++c) { double t = a[c] - b[c]; s += t * t; }
    return sqrt(s > 1e-9 ? s : 1e-9);
  }
  double na = 0.0, nb = 0.0, ab = 0.0;
  for (int c = 0; c < dim; ++c) { na += a[c] * a[c]; nb += b[c] * b[c]; }
  na = sqrt(na); nb = sqrt(nb);
  const double ca = na > 1e-8 ? na : 1e-8, cb = nb > 1e-8 ? nb : 1e-8;
  for (int c = 0; c < dim; ++c) ab += (a[c] / ca) * (b[c] / cb);
  return 1.0 - ab;
}

/* d(dist)/da (gradient w.r.t. the first argument); by symmetry use with swapped args for b */
static void pair_dist_grad(const double *a, const double *b, int dim, int metric, double *ga) {
  if (metric == 0) {
    double s = 0.0;
    for (int c = 0; c < dim; ++c) { double t = a[c] - b[c]; s += t * t; }
    if (s > 1e-9) {
      const double r = sqrt(s);
      for (int c = 0; c < dim; ++c) ga[c] = (a[c] - b[c]) / r;
    } else {
      for (int c = 0; c < dim; ++c) ga[c] = 0.0; /* clamped branch of torch.maximum */
    }
    return;
  }
  /* cosine: the clamp in ATen's cosine_similarity is applied under no_grad, so the
   * gradient is that of (a/|a|_c).(b/|b|_c) with |a|_c treated as |a| whenever
   * unclamped; for clamped norms (|a| < 1e-8) the norm's own gradient still flows. */
  double na = 0.0, nb = 0.0;
  for (int c = 0; c < dim; ++c) { na += a[c] * a[c]; nb += b[c] * b[c]; }
  na = sqrt(na); nb = sqrt(nb);
  const double ca = na > 1e-8 ? na : 1e-8, cb = nb > 1e-8 ? nb : 1e-8;
  double dot = 0.0;
  for (int c = 0; c < dim; ++c) dot += a[c] * b[c];
  for (int c = 0; c < dim; ++c) {
    /* cos = dot/(ca*cb); d cos/da = b/(ca cb) - dot/(ca^2 cb) * d|a|/da, d|a|/da = a/|a| */
    double dn = na > 0.0 ? a[c] / na : 0.0;
    double dcos = b[c] / (ca * cb) - dot / (ca * ca * cb) * dn;
    ga[c] = -dcos;
  }
}

double vv_oracle_quartet_pass_f64(const double *x, const double *z, int64_t B, int D, int L,
                                  const int64_t *perm, int metric_hd, int metric_ld,
                                  double scale_grad, double *grad_z) {
  const int64_t nq = B / 4;
  if (nq == 0) return NAN; /* mean over an empty set, as in the reference */
  double total = 0.0;
  double *g = (double *)malloc((size_t)L * sizeof(double));
  for (int64_t q = 0; q < nq; ++q) {
    int64_t r[4];
    for (int a = 0; a < 4; ++a) r[a] = perm ? perm[a * nq + q] : a * nq + q;
    double dx[6], dz[6], X = 0.0, S = 0.0;
    for (int p = 0; p < 6; ++p) {
      dx[p] = pair_dist(x + r[VV_PA[p]] * D, x + r[VV_PB[p]] * D, D, metric_hd);
      dz[p] = pair_dist(z + r[VV_PA[p]] * L, z + r[VV_PB[p]] * L, L, metric_ld);
      X += dx[p]; S += dz[p];
    }
    double e[6], cost = 0.0, er = 0.0;
    for (int p = 0; p < 6; ++p) {
      e[p] = dz[p] / S - dx[p] / X;
      cost += e[p] * e[p];
      er += e[p] * (dz[p] / S);
    }
    total += cost;
    if (grad_z) {
      for (int p = 0; p < 6; ++p) {
        const double w = (2.0 / S) * (e[p] - er) * scale_grad; /* dC/d(dz_p) */
        const int64_t ra = r[VV_PA[p]], rb = r[VV_PB[p]];
        pair_dist_grad(z + ra * L, z + rb * L, L, metric_ld, g);
        for (int c = 0; c < L; ++c) grad_z[ra * L + c] += w * g[c];
        pair_dist_grad(z + rb * L, z + ra * L, L, metric_ld, g);
        for (int c = 0; c < L; ++c) grad_z[rb * L + c] += w * g[c];
      }
    }
  }
  free(g);
  return total / (double)nq * (double)B / (double)nq;
}
