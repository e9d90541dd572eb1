// extern "C" surface of libvivae_b200.so (declared in include/vivae_b200.h).
#include <cmath>
#include <cstring>
#include <string>

#include "common.cuh"

namespace vvb {

// ---------------------------------------------------------------- shared plumbing
static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// launchers implemented in the kernel translation units
size_t brute_workspace_bytes(int64_t nq, int k);
int launch_knn_brute(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev, int64_t q0,
                     int64_t nq, int64_t *out_idx, float *out_dist, int64_t out_row0, void *workspace,
                     size_t workspace_bytes, cudaStream_t stream);
size_t knn_tensor_workspace_bytes(int64_t n, int d, int k, int64_t nq);
bool knn_tensor_supported(int64_t n, int d, int k);
int launch_knn_tensor(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq, int64_t *out_idx,
                      float *out_dist, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                      vvb_knn_stats *stats);
int knn_screen_debug(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq, uint32_t *cand_idx_out,
                     float *cand_key_out, int *cand_cnt_out, float *cand_thr_out, float *xnorm_out,
                     float *scale_ymax_out, float *mu_out);
int knn_screen_debug_capacity();
size_t smooth_workspace_bytes(int elem_size, int64_t n, int d, int n_iter);
int launch_smooth(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                  double coef, int n_iter, int mode, void *out, void *workspace, size_t workspace_bytes,
                  cudaStream_t st);
size_t quartet_workspace_bytes(int64_t B);
int launch_quartet_fwd(const float *x, const float *z, int64_t B, int D, int L, const int64_t *const *perms,
                       int n_sampling, int mhd, int mld, float *loss_out, float *grad_unit, void *workspace,
                       cudaStream_t st);
int launch_quartet_bwd(const float *grad_unit, const float *grad_out, int64_t count, float *grad_z,
                       cudaStream_t st);

struct DevBuf {  // RAII device allocation for the *_host entry points
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (bytes == 0) bytes = 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      cudaGetLastError();
      return e == cudaErrorMemoryAllocation ? VVB_ENOMEM : VVB_ECUDA;
    }
    return VVB_OK;
  }
};

struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  explicit EventTimer(cudaStream_t s) : st(s) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~EventTimer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void start() { cudaEventRecord(a, st); }
  void stop() { cudaEventRecord(b, st); }
  float ms() {
    float v = 0.f;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&v, a, b);
    return v;
  }
};

static int knn_check_args(const void *x, int64_t n, int d, int k, int64_t q0, int64_t nq, const void *idx,
                          const void *dist) {
  VVB_REQUIRE(x && idx && dist, "make_knn: null pointer argument");
  VVB_REQUIRE(n >= 2 && d >= 1, "make_knn: need at least 2 rows and 1 column");
  VVB_REQUIRE(k >= 1 && k <= n - 1, "`k` must be between 1 and (n-1)");  // knn.py:119-120
  VVB_REQUIRE(q0 >= 0 && nq >= 0 && q0 + nq <= n, "make_knn: query range [%lld, %lld) outside [0, %lld)",
              (long long)q0, (long long)(q0 + nq), (long long)n);
  return VVB_OK;
}

static int resolve_algo(int algo, int64_t n, int d, int k, int *out) {
  VVB_REQUIRE(algo == VVB_KNN_AUTO || algo == VVB_KNN_BRUTE || algo == VVB_KNN_TENSOR, "make_knn: unknown algo %d", algo);
  const bool ok = knn_tensor_supported(n, d, k);
  if (algo == VVB_KNN_TENSOR && !ok) {
    set_error("make_knn: tensor-core path does not support n=%lld d=%d k=%d", (long long)n, d, k);
    return VVB_ENOTSUP;
  }
  *out = (algo == VVB_KNN_BRUTE || !ok) ? VVB_KNN_BRUTE : VVB_KNN_TENSOR;
  return VVB_OK;
}

}  // namespace vvb

using namespace vvb;

extern "C" {

int vvb_version(void) { return 100; }
const char *vvb_last_error(void) { return g_error.c_str(); }
int64_t vvb_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------ kNN
int vvb_make_knn_workspace_bytes(int64_t n, int d, int k, int64_t nq, int algo, size_t *bytes) {
  VVB_REQUIRE(bytes != nullptr, "make_knn: null bytes pointer");
  VVB_REQUIRE(n >= 2 && d >= 1 && k >= 1 && k <= n - 1 && nq >= 0, "make_knn: bad shape arguments");
  int use = 0;
  int rc = resolve_algo(algo, n, d, k, &use);
  if (rc) return rc;
  *bytes = (use == VVB_KNN_BRUTE) ? brute_workspace_bytes(nq, k) : knn_tensor_workspace_bytes(n, d, k, nq);
  return VVB_OK;
}

int vvb_make_knn_device(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq, int algo,
                        int64_t *idx_out_dev, float *dist_out_dev, void *workspace_dev, size_t workspace_bytes,
                        void *stream, vvb_knn_stats *stats) {
  int rc = knn_check_args(x_dev, n, d, k, q0, nq, idx_out_dev, dist_out_dev);
  if (rc) return rc;
  int use = 0;
  rc = resolve_algo(algo, n, d, k, &use);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t launches0 = g_launches.load();
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rows = nq;
    stats->algo_used = use;
  }
  if (use == VVB_KNN_BRUTE) {
    EventTimer tm(st);
    if (stats) tm.start();
    rc = launch_knn_brute(x_dev, n, d, k, nullptr, q0, nq, idx_out_dev, dist_out_dev, q0, workspace_dev,
                          workspace_bytes, st);
    if (rc) return rc;
    if (stats) {
      tm.stop();
      stats->ms_fallback = tm.ms();
      stats->rows_fallback = nq;
    }
  } else {
    rc = launch_knn_tensor(x_dev, n, d, k, q0, nq, idx_out_dev, dist_out_dev, workspace_dev, workspace_bytes, st,
                           stats);
    if (rc) return rc;
  }
  if (stats) {
    VVB_CUDA(cudaStreamSynchronize(st));
    stats->kernel_launches = g_launches.load() - launches0;
  }
  return VVB_OK;
}

int vvb_make_knn_host(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq, int algo,
                      int64_t *idx_out_host, float *dist_out_host, vvb_knn_stats *stats) {
  int rc = knn_check_args(x_host, n, d, k, q0, nq, idx_out_host, dist_out_host);
  if (rc) return rc;
  size_t ws_bytes = 0;
  rc = vvb_make_knn_workspace_bytes(n, d, k, nq, algo, &ws_bytes);
  if (rc) return rc;
  DevBuf x, idx, dist, ws;
  if ((rc = x.alloc(static_cast<size_t>(n) * d * sizeof(float)))) return rc;
  if ((rc = idx.alloc(static_cast<size_t>(nq) * (k + 1) * sizeof(int64_t)))) return rc;
  if ((rc = dist.alloc(static_cast<size_t>(nq) * (k + 1) * sizeof(float)))) return rc;
  if ((rc = ws.alloc(ws_bytes))) return rc;
  cudaStream_t st = nullptr;
  VVB_CUDA(cudaMemcpyAsync(x.p, x_host, static_cast<size_t>(n) * d * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = vvb_make_knn_device(static_cast<const float *>(x.p), n, d, k, q0, nq, algo, static_cast<int64_t *>(idx.p),
                           static_cast<float *>(dist.p), ws.p, ws_bytes, st, stats);
  if (rc) return rc;
  VVB_CUDA(cudaMemcpyAsync(idx_out_host, idx.p, static_cast<size_t>(nq) * (k + 1) * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, st));
  VVB_CUDA(cudaMemcpyAsync(dist_out_host, dist.p, static_cast<size_t>(nq) * (k + 1) * sizeof(float),
                           cudaMemcpyDeviceToHost, st));
  VVB_CUDA(cudaStreamSynchronize(st));
  return VVB_OK;
}

int vvb_knn_screen_debug_capacity(void) { return knn_screen_debug_capacity(); }

int vvb_knn_screen_debug_host(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq,
                              uint32_t *cand_idx_out, float *cand_key_out, int32_t *cand_cnt_out,
                              float *cand_thr_out, float *xnorm_out, float *scale_ymax_out, float *mu_out) {
  VVB_REQUIRE(x_host && cand_idx_out && cand_key_out && cand_cnt_out && cand_thr_out && xnorm_out && scale_ymax_out &&
                  mu_out, "screen debug: null pointer argument");
  return knn_screen_debug(x_host, n, d, k, q0, nq, cand_idx_out, cand_key_out, cand_cnt_out, cand_thr_out, xnorm_out,
                          scale_ymax_out, mu_out);
}

// --------------------------------------------------------------------- smooth
int vvb_smooth_workspace_bytes(int elem_size, int64_t n, int d, int n_iter, size_t *bytes) {
  VVB_REQUIRE(bytes != nullptr, "smooth: null bytes pointer");
  VVB_REQUIRE((elem_size == 4 || elem_size == 8) && n >= 1 && d >= 1 && n_iter >= 1, "smooth: bad shape arguments");
  *bytes = smooth_workspace_bytes(elem_size, n, d, n_iter);
  return VVB_OK;
}

int vvb_smooth_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_dev, int64_t ld_idx,
                      int k, double coef, int n_iter, int mode, void *out_dev, void *workspace_dev,
                      size_t workspace_bytes, void *stream) {
  VVB_REQUIRE(x_dev && idx_dev && out_dev && workspace_dev, "smooth: null pointer argument");
  // the reference's own argument checks (knn.py:165-170), restated at the ABI
  VVB_REQUIRE(k >= 1 && k <= n - 1, "`k` must be between 1 and (n-1)");
  VVB_REQUIRE(coef >= 0.0 && coef <= 1.0, "`coef` must be more than 0. and at most 1.");
  VVB_REQUIRE(n_iter >= 1 && n_iter <= 10000, "`n_iter` must be at least 1 and at most 10000");
  return launch_smooth(x_dev, elem_size, n, d, idx_dev, ld_idx, k, coef, n_iter, mode, out_dev, workspace_dev,
                       workspace_bytes, static_cast<cudaStream_t>(stream));
}

int vvb_smooth_host(const void *x_host, int elem_size, int64_t n, int d, const int64_t *idx_host, int64_t ld_idx,
                    int k, double coef, int n_iter, int mode, void *out_host) {
  VVB_REQUIRE(x_host && idx_host && out_host, "smooth: null pointer argument");
  size_t ws_bytes = 0;
  int rc = vvb_smooth_workspace_bytes(elem_size, n, d, n_iter, &ws_bytes);
  if (rc) return rc;
  VVB_REQUIRE(ld_idx >= k && k >= 1, "smooth: neighbour matrix has fewer than k columns");
  const size_t xb = static_cast<size_t>(n) * d * elem_size;
  DevBuf x, out, idx, ws;
  if ((rc = x.alloc(xb))) return rc;
  if ((rc = out.alloc(xb))) return rc;
  // only the first k columns are consumed: ship a compact (n, k) copy
  if ((rc = idx.alloc(static_cast<size_t>(n) * k * sizeof(int64_t)))) return rc;
  if ((rc = ws.alloc(ws_bytes))) return rc;
  cudaStream_t st = nullptr;
  VVB_CUDA(cudaMemcpyAsync(x.p, x_host, xb, cudaMemcpyHostToDevice, st));
  VVB_CUDA(cudaMemcpy2DAsync(idx.p, static_cast<size_t>(k) * sizeof(int64_t), idx_host,
                             static_cast<size_t>(ld_idx) * sizeof(int64_t), static_cast<size_t>(k) * sizeof(int64_t),
                             static_cast<size_t>(n), cudaMemcpyHostToDevice, st));
  rc = vvb_smooth_device(x.p, elem_size, n, d, static_cast<const int64_t *>(idx.p), k, k, coef, n_iter, mode, out.p,
                         ws.p, ws_bytes, st);
  if (rc) return rc;
  VVB_CUDA(cudaMemcpyAsync(out_host, out.p, xb, cudaMemcpyDeviceToHost, st));
  VVB_CUDA(cudaStreamSynchronize(st));
  return VVB_OK;
}

// -------------------------------------------------------------------- quartet
size_t vvb_quartet_workspace_bytes(int64_t B) { return quartet_workspace_bytes(B < 0 ? 0 : B); }

int vvb_quartet_loss_fwd_device(const float *x_dev, const float *z_dev, int64_t B, int D, int L,
                                const int64_t *const *perms, int n_sampling, int metric_hd, int metric_ld,
                                float *loss_out_dev, float *grad_unit_dev, void *workspace_dev, void *stream) {
  VVB_REQUIRE(x_dev && z_dev && loss_out_dev && workspace_dev, "quartet loss: null pointer argument");
  return launch_quartet_fwd(x_dev, z_dev, B, D, L, perms, n_sampling, metric_hd, metric_ld, loss_out_dev,
                            grad_unit_dev, workspace_dev, static_cast<cudaStream_t>(stream));
}

int vvb_quartet_loss_bwd_device(const float *grad_unit_dev, const float *grad_out_dev, int64_t count,
                                float *grad_z_dev, void *stream) {
  VVB_REQUIRE(grad_unit_dev && grad_out_dev && grad_z_dev, "quartet loss backward: null pointer argument");
  return launch_quartet_bwd(grad_unit_dev, grad_out_dev, count, grad_z_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
