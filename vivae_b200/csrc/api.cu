// extern "C" surface of libvivae_b200.so (declared in include/vivae_b200.h).
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

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
size_t knn_tensor_workspace_bytes(int64_t n, int d, int k, int64_t nq, bool host_path);
bool knn_tensor_supported(int64_t n, int d, int k);
int64_t knn_tensor_chunk_rows(int64_t nq, bool host_path);
size_t knn_tensor_session_bytes();
int knn_tensor_begin(void *session_mem, const float *x_dev, int64_t n, int d, int k, int64_t chunk_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st, vvb_knn_stats *stats, bool planned);
int knn_tensor_plan_phase(const float *x_dev, int64_t n, int d, int k, int64_t chunk_rows, void *workspace,
                          size_t workspace_bytes, int phase, int part, int parts, cudaStream_t st, vvb_exchange *xch,
                          int *n_xch);
int knn_tensor_chunk(void *session_mem, int64_t q0, int64_t nq, int64_t *out_idx, float *out_dist, cudaStream_t st,
                     vvb_knn_stats *stats, bool screen_only);
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
int smooth_begin(void *workspace, int elem_size, int64_t n, int d, int n_iter, cudaStream_t st);
int launch_smooth_rows(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                       double coef, int mode, void *out, void *workspace, size_t workspace_bytes, int chunk,
                       int64_t row_lo, int64_t row_hi, cudaStream_t st);
int smooth_max_chunks();
int launch_smooth_cols(const void *x, int elem_size, int64_t n, int d, int col0, int dcols, const int64_t *idx,
                       int64_t ld_idx, int k, double coef, int n_iter, int mode, void *out, void *workspace,
                       size_t workspace_bytes, cudaStream_t st);
int launch_knn_merge(const float *xc, int64_t nc, int d, int64_t q0, int64_t nq, const int64_t *new_idx, int kn,
                     int64_t keep_lo, int64_t keep_hi, int64_t global_off, int64_t *run_idx, float *run_d2, int k,
                     cudaStream_t st);
int launch_knn_merge_init(int64_t *run_idx, float *run_d2, int64_t nq, int k, int64_t self0, cudaStream_t st);
int launch_knn_merge_finish(const float *run_d2, int64_t count, float *dist, cudaStream_t st);
int launch_quartet_bwd(const float *grad_unit, const float *grad_out, int64_t count, float *grad_z,
                       cudaStream_t st);

int launch_vae_step(const float *const *w_ptrs, const float *const *b_ptrs, const int *in_dims, const int *out_dims,
                    int n_enc, int n_dec, const float *x, const float *eps, int B, int D, int L, float lam_recon,
                    float lam_kldiv, float lam_mds, float *gpart, float *lpart, float *gflat, float *sums,
                    float *loss_out, size_t *smem_bytes, int *n_params, int *n_cta, cudaStream_t st);

size_t smooth_peer_workspace_bytes();
int launch_smooth_peer(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                       double coef, int epoch, int rank, int world, void *const *ll_ptrs, void *workspace,
                       cudaStream_t st);
int launch_smooth_peer_unpack(const void *ll_local, int elem_size, int64_t count, int epoch, void *out, void *workspace,
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

// True if [p, p + bytes) is page-locked memory CUDA knows about (cudaHostAlloc / cudaHostRegister): the arrays
// this package returns come from such a pool (vvb_host_alloc), and callers may pass their own pinned buffers.
static bool host_range_is_pinned(const void *p, size_t bytes) {
  if (!p || bytes == 0) return false;
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || a0.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return false;
  }
  if (cudaPointerGetAttributes(&a1, static_cast<const char *>(p) + bytes - 1) != cudaSuccess ||
      a1.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// Touch the pages of freshly allocated host output arrays on helper threads.
struct HostPrefault {
  struct Range {
    char *p;
    size_t bytes;
  };
  std::vector<Range> ranges;
  std::vector<std::thread> threads;
  void add(void *p, size_t bytes) {
    if (p && bytes >= (8u << 20) && !host_range_is_pinned(p, bytes)) ranges.push_back({static_cast<char *>(p), bytes});
  }
  void start() {
    const size_t page = static_cast<size_t>(sysconf(_SC_PAGESIZE));
    for (const Range &r : ranges) {
      // whole pages inside the range only: bytes outside the array are never written
      uintptr_t lo = (reinterpret_cast<uintptr_t>(r.p) + page - 1) / page * page;
      uintptr_t hi = (reinterpret_cast<uintptr_t>(r.p) + r.bytes) / page * page;
      if (hi <= lo) continue;
      madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);  // best effort (THP = madvise)
      const int nthreads = 4;
      const size_t pages = (hi - lo) / page;
      for (int t = 0; t < nthreads; ++t) {
        const size_t p0 = pages * t / nthreads, p1 = pages * (t + 1) / nthreads;
        threads.emplace_back([lo, page, p0, p1]() {
          volatile char *base = reinterpret_cast<volatile char *>(lo);
          for (size_t i = p0; i < p1; ++i) base[i * page] = 0;  // the array is an OUTPUT: contents are overwritten
        });
      }
    }
  }
  void join() {
    for (auto &t : threads) t.join();
    threads.clear();
  }
  ~HostPrefault() { join(); }
};

// Host <-> device copies of large PAGEABLE arrays (the drop-in API takes and returns NumPy arrays).
// A plain cudaMemcpy from pageable memory is staged by the driver through one bounce buffer by one thread
// (~10 GB/s measured on the bench host).  Here T worker threads each own a slice of the array, two pinned
// bounce buffers and a stream: memcpy into pinned memory and DMA out of it overlap, and the T memcpys run in
// parallel.  Blocking: returns when the data has arrived.
// Neighbour indices cross PCIe as int32 and live as int64 on both sides (the reference's dtype, knn.py:137):
// the staging memcpy is where they are widened / narrowed, so the conversion costs no extra pass and the
// copies -- bound by the host's memory bandwidth, not by PCIe -- move a third less.
enum CopyConv { CONV_NONE = 0, CONV_WIDEN_I32 = 1 /* D2H: device int32 -> host int64 */,
                CONV_NARROW_I64 = 2 /* H2D: host int64 -> device int32, out-of-range -> -1 */ };

static void widen_i32(int64_t *dst, const int32_t *src, size_t count) {
  for (size_t i = 0; i < count; ++i) dst[i] = src[i];
}
static void narrow_i64(int32_t *dst, const int64_t *src, size_t count) {
  for (size_t i = 0; i < count; ++i) {
    const int64_t v = src[i];
    dst[i] = (v >= 0 && v <= INT32_MAX) ? static_cast<int32_t>(v) : -1;  // -1 is flagged by the kernel
  }
}

struct StagedCopier {
  static constexpr int T = 16;  // bounce-buffer sets; threads() of them are used
  static constexpr size_t CH_MAX = 4u << 20;
  static size_t chunk() {  // bytes per bounce buffer actually used (VVB_COPY_CHUNK_KB, multiple of 4 KB, <= 4 MB)
    static size_t c = [] {
      const char *e = getenv("VVB_COPY_CHUNK_KB");
      size_t v = (e ? static_cast<size_t>(atoi(e)) : 4096) << 10;
      v = v / 4096 * 4096;
      return v < 4096 ? static_cast<size_t>(4096) : (v > CH_MAX ? CH_MAX : v);
    }();
    return c;
  }
  static int threads() {
    static int t = [] {
      const char *e = getenv("VVB_COPY_THREADS");
      const int v = e ? atoi(e) : 6;  // measured on the bench host: see DESIGN.md section 5
      return v < 1 ? 1 : (v > T ? T : v);
    }();
    return t;
  }
  char *pin[T][2] = {};
  cudaStream_t st[T] = {};
  cudaEvent_t ev[T][2] = {};
  bool ready = false;
  int init() {
    if (ready) return VVB_OK;
    for (int t = 0; t < T; ++t) {
      VVB_CUDA(cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        VVB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&pin[t][b]), CH_MAX, cudaHostAllocDefault));
        VVB_CUDA(cudaEventCreateWithFlags(&ev[t][b], cudaEventDisableTiming));
      }
    }
    ready = true;
    return VVB_OK;
  }
  void release() {
    if (!ready) return;
    for (int t = 0; t < T; ++t) {
      for (int b = 0; b < 2; ++b) {
        cudaFreeHost(pin[t][b]);
        cudaEventDestroy(ev[t][b]);
        pin[t][b] = nullptr;
      }
      cudaStreamDestroy(st[t]);
    }
    ready = false;
  }
  // dir = cudaMemcpyHostToDevice: dst device, src host;  cudaMemcpyDeviceToHost: dst host, src device.
  // `after`: event the device side must wait for before it is read (D2H) or overwritten (H2D); may be null.
  // `bytes` counts the DEVICE side; with a conversion the host side is twice as large.
  int copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind dir, cudaEvent_t after, int conv = CONV_NONE) {
    if (bytes == 0) return VVB_OK;
    int rc = init();
    if (rc) return rc;
    int dev = 0;
    VVB_CUDA(cudaGetDevice(&dev));
    if (conv == CONV_NONE &&
        host_range_is_pinned(dir == cudaMemcpyHostToDevice ? src : static_cast<const void *>(dst), bytes)) {
      // page-locked host memory: one DMA at PCIe speed, no bounce buffers, no CPU pass
      if (after) VVB_CUDA(cudaStreamWaitEvent(st[0], after, 0));
      VVB_CUDA(cudaMemcpyAsync(dst, src, bytes, dir, st[0]));
      VVB_CUDA(cudaStreamSynchronize(st[0]));
      return VVB_OK;
    }
    const size_t CH = chunk();
    const size_t nch = (bytes + CH - 1) / CH;
    const int nt = static_cast<int>(nch < static_cast<size_t>(threads()) ? nch : threads());
    std::atomic<int> failed{0};
    auto work = [&](int t) {
      cudaSetDevice(dev);
      if (after) cudaStreamWaitEvent(st[t], after, 0);
      const size_t c0 = nch * t / nt, c1 = nch * (t + 1) / nt;
      char *d = static_cast<char *>(dst);
      const char *s = static_cast<const char *>(src);
      if (dir == cudaMemcpyHostToDevice) {
        for (size_t c = c0; c < c1; ++c) {
          const int b = static_cast<int>((c - c0) & 1);
          const size_t off = c * CH, len = (off + CH <= bytes) ? CH : bytes - off;
          if (c - c0 >= 2 && cudaEventSynchronize(ev[t][b]) != cudaSuccess) failed = 1;  // bounce buffer free again
          if (conv == CONV_NARROW_I64)
            narrow_i64(reinterpret_cast<int32_t *>(pin[t][b]), reinterpret_cast<const int64_t *>(s + 2 * off), len / 4);
          else
            memcpy(pin[t][b], s + off, len);
          if (cudaMemcpyAsync(d + off, pin[t][b], len, cudaMemcpyHostToDevice, st[t]) != cudaSuccess) failed = 1;
          cudaEventRecord(ev[t][b], st[t]);
        }
      } else {
        // DMA of chunk c overlaps the memcpy of chunk c-1 out of the other bounce buffer
        for (size_t c = c0; c <= c1; ++c) {
          if (c < c1) {
            const int b = static_cast<int>((c - c0) & 1);
            const size_t off = c * CH, len = (off + CH <= bytes) ? CH : bytes - off;
            if (cudaMemcpyAsync(pin[t][b], s + off, len, cudaMemcpyDeviceToHost, st[t]) != cudaSuccess) failed = 1;
            cudaEventRecord(ev[t][b], st[t]);
          }
          if (c > c0) {
            const int b = static_cast<int>((c - 1 - c0) & 1);
            const size_t off = (c - 1) * CH, len = (off + CH <= bytes) ? CH : bytes - off;
            if (cudaEventSynchronize(ev[t][b]) != cudaSuccess) failed = 1;
            if (conv == CONV_WIDEN_I32)
              widen_i32(reinterpret_cast<int64_t *>(d + 2 * off), reinterpret_cast<const int32_t *>(pin[t][b]), len / 4);
            else
              memcpy(d + off, pin[t][b], len);
          }
        }
      }
      if (cudaStreamSynchronize(st[t]) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed) {
      set_error("staged host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      return VVB_ECUDA;
    }
    return VVB_OK;
  }
};

// Grow-only device buffers and two streams kept between *_host calls (a 2 GB cudaMalloc/cudaFree
// pair costs ~27 ms per call otherwise).  Released by vvb_release_cached_memory().
struct DeviceCache {
  std::mutex mu;
  int device = -1;
  void *ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  StagedCopier copier;
  StagedCopier copier2;  // second set of bounce buffers: uploads and downloads that run at the same time
  int last_rc = VVB_OK;
  static DeviceCache &get() {
    static DeviceCache c;
    return c;
  }
  void release() {
    for (int i = 0; i < 8; ++i) {
      if (ptr[i]) cudaFree(ptr[i]);
      ptr[i] = nullptr;
      cap[i] = 0;
    }
    if (s_compute) cudaStreamDestroy(s_compute);
    if (s_copy) cudaStreamDestroy(s_copy);
    s_compute = s_copy = nullptr;
    copier.release();
    copier2.release();
  }
  void check_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (dev != device) {
      if (device >= 0) {
        cudaSetDevice(device);
        release();
        cudaSetDevice(dev);
      }
      device = dev;
    }
  }
  // Slots handed out since begin_request(): a failed allocation may only evict the OTHER slots -- the caller still
  // holds raw pointers into these (and may have DMAs in flight on the streams, which are never destroyed here).
  unsigned in_use = 0;
  void begin_request() {
    check_device();
    in_use = 0;
    last_rc = VVB_OK;
  }
  void *slot(int i, size_t bytes) {
    check_device();
    if (bytes == 0) bytes = 256;
    in_use |= 1u << i;
    if (cap[i] >= bytes) return ptr[i];
    if (ptr[i]) cudaFree(ptr[i]);
    ptr[i] = nullptr;
    cap[i] = 0;
    cudaError_t e = cudaMalloc(&ptr[i], bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // make room: drop the cached buffers this request has not been given, then retry once
      for (int j = 0; j < 8; ++j) {
        if ((in_use >> j) & 1u) continue;
        if (ptr[j]) cudaFree(ptr[j]);
        ptr[j] = nullptr;
        cap[j] = 0;
      }
      e = cudaMalloc(&ptr[i], bytes);
    }
    if (e != cudaSuccess) {
      ptr[i] = nullptr;
      set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      cudaGetLastError();
      last_rc = (e == cudaErrorMemoryAllocation) ? VVB_ENOMEM : VVB_ECUDA;
      return nullptr;
    }
    cap[i] = bytes;
    return ptr[i];
  }
  cudaStream_t compute_stream() {
    check_device();
    if (!s_compute && cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      s_compute = nullptr;
    }
    return s_compute;
  }
  cudaStream_t copy_stream() {
    check_device();
    if (!s_copy && cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      s_copy = nullptr;
    }
    return s_copy;
  }
};

__global__ void narrow_idx_kernel(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) out[i] = static_cast<int32_t>(in[i]);  // row indices (or -1) of a matrix with n <= INT32_MAX rows
}
__global__ void widen_idx_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) out[i] = in[i];
}

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
  *bytes = (use == VVB_KNN_BRUTE) ? brute_workspace_bytes(knn_tensor_chunk_rows(nq, false), k)
                                   : knn_tensor_workspace_bytes(n, d, k, nq, false);
  return VVB_OK;
}

}  // extern "C"

// One pass of the graph builder over query rows [q0, q0+nq) in chunks; `after_chunk(c0, rows)` is
// called once the work of a chunk has been ENQUEUED (used by the host entry point to overlap the
// device-to-host copy of chunk i with the computation of chunk i+1).
template <typename Hook>
static int knn_run(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq, int use, int64_t *idx_dev,
                   float *dist_dev, void *ws, size_t ws_bytes, cudaStream_t st, vvb_knn_stats *stats, bool host_path,
                   Hook after_chunk, bool planned = false) {
  const int64_t chunk = knn_tensor_chunk_rows(nq, host_path);
  std::vector<char> session(knn_tensor_session_bytes() + 64);
  void *sess = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(session.data()) + 63) & ~uintptr_t(63));
  int rc;
  if (use == VVB_KNN_TENSOR) {
    rc = knn_tensor_begin(sess, x_dev, n, d, k, chunk, ws, ws_bytes, st, stats, planned);
    if (rc) return rc;
  }
  for (int64_t c0 = 0; c0 < nq; c0 += chunk) {
    const int64_t rows = std::min(chunk, nq - c0);
    int64_t *oi = idx_dev + c0 * (k + 1);
    float *od = dist_dev + c0 * (k + 1);
    if (use == VVB_KNN_TENSOR) {
      rc = knn_tensor_chunk(sess, q0 + c0, rows, oi, od, st, stats, false);
    } else {
      EventTimer tm(st);
      if (stats) tm.start();
      rc = launch_knn_brute(x_dev, n, d, k, nullptr, q0 + c0, rows, oi, od, q0 + c0, ws, ws_bytes, st);
      if (!rc && stats) {
        tm.stop();
        stats->ms_fallback += tm.ms();
        stats->rows_fallback += rows;
      }
    }
    if (rc) return rc;
    rc = after_chunk(c0, rows);
    if (rc) return rc;
  }
  return VVB_OK;
}

static size_t knn_ws_bytes(int use, int64_t n, int d, int k, int64_t nq, bool host_path) {
  return (use == VVB_KNN_BRUTE) ? brute_workspace_bytes(knn_tensor_chunk_rows(nq, host_path), k)
                                : knn_tensor_workspace_bytes(n, d, k, nq, host_path);
}

extern "C" {

int vvb_make_knn_device(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq, int algo,
                        int64_t *idx_out_dev, float *dist_out_dev, void *workspace_dev, size_t workspace_bytes,
                        void *stream, vvb_knn_stats *stats) {
  int rc = knn_check_args(x_dev, n, d, k, q0, nq, idx_out_dev, dist_out_dev);
  if (rc) return rc;
  int use = 0;
  rc = resolve_algo(algo, n, d, k, &use);
  if (rc) return rc;
  VVB_REQUIRE(workspace_bytes >= knn_ws_bytes(use, n, d, k, nq, false), "make_knn: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t launches0 = g_launches.load();
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rows = nq;
    stats->algo_used = use;
  }
  rc = knn_run(x_dev, n, d, k, q0, nq, use, idx_out_dev, dist_out_dev, workspace_dev, workspace_bytes, st, stats,
               /*host_path=*/false, [](int64_t, int64_t) { return VVB_OK; });
  if (rc) return rc;
  if (stats) {
    VVB_CUDA(cudaStreamSynchronize(st));
    stats->kernel_launches = g_launches.load() - launches0;
  }
  return VVB_OK;
}

int vvb_knn_plan_phase_device(const float *x_dev, int64_t n, int d, int k, int64_t nq, void *workspace_dev,
                              size_t workspace_bytes, int phase, int part, int parts, void *stream, vvb_exchange *xch,
                              int *n_xch) {
  VVB_REQUIRE(x_dev && workspace_dev && xch && n_xch, "knn plan: null pointer argument");
  VVB_REQUIRE(n >= 2 && d >= 1 && k >= 1 && k <= n - 1 && nq >= 0, "knn plan: bad shape arguments");
  if (!knn_tensor_supported(n, d, k)) {
    set_error("knn plan: n=%lld d=%d k=%d is outside the tensor-core path", (long long)n, d, k);
    return VVB_ENOTSUP;
  }
  return knn_tensor_plan_phase(x_dev, n, d, k, knn_tensor_chunk_rows(nq, false), workspace_dev, workspace_bytes, phase,
                               part, parts, static_cast<cudaStream_t>(stream), xch, n_xch);
}

int vvb_make_knn_planned_device(const float *x_dev, int64_t n, int d, int k, int64_t q0, int64_t nq,
                                int64_t *idx_out_dev, float *dist_out_dev, void *workspace_dev, size_t workspace_bytes,
                                void *stream, vvb_knn_stats *stats) {
  int rc = knn_check_args(x_dev, n, d, k, q0, nq, idx_out_dev, dist_out_dev);
  if (rc) return rc;
  if (!knn_tensor_supported(n, d, k)) {
    set_error("make_knn (planned): n=%lld d=%d k=%d is outside the tensor-core path", (long long)n, d, k);
    return VVB_ENOTSUP;
  }
  VVB_REQUIRE(workspace_bytes >= knn_ws_bytes(VVB_KNN_TENSOR, n, d, k, nq, false), "make_knn: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t launches0 = g_launches.load();
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rows = nq;
    stats->algo_used = VVB_KNN_TENSOR;
  }
  rc = knn_run(x_dev, n, d, k, q0, nq, VVB_KNN_TENSOR, idx_out_dev, dist_out_dev, workspace_dev, workspace_bytes, st,
               stats, /*host_path=*/false, [](int64_t, int64_t) { return VVB_OK; }, /*planned=*/true);
  if (rc) return rc;
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
  int use = 0;
  rc = resolve_algo(algo, n, d, k, &use);
  if (rc) return rc;
  const size_t ws_bytes = knn_ws_bytes(use, n, d, k, nq, true);
  const size_t idx_bytes = static_cast<size_t>(nq) * (k + 1) * sizeof(int64_t);
  const size_t dist_bytes = static_cast<size_t>(nq) * (k + 1) * sizeof(float);
  // the output arrays are usually freshly allocated: fault their pages in on helper threads while
  // the GPU works (a D2H copy into untouched pageable memory runs 8x slower than into touched pages)
  HostPrefault pf;
  pf.add(idx_out_host, idx_bytes);
  pf.add(dist_out_host, dist_bytes);
  pf.start();
  DeviceCache &dc = DeviceCache::get();
  std::lock_guard<std::mutex> lock(dc.mu);
  dc.begin_request();
  float *x_dev = static_cast<float *>(dc.slot(0, static_cast<size_t>(n) * d * sizeof(float)));
  int64_t *idx_dev = static_cast<int64_t *>(dc.slot(1, idx_bytes));
  float *dist_dev = static_cast<float *>(dc.slot(2, dist_bytes));
  void *ws = dc.slot(3, ws_bytes);
  // indices travel as int32 (see CopyConv) unless the output array is page-locked: then the DMA writes it directly
  const bool narrow = n <= INT32_MAX && !host_range_is_pinned(idx_out_host, idx_bytes);
  int32_t *idx32_dev = narrow ? static_cast<int32_t *>(dc.slot(5, idx_bytes / 2)) : nullptr;
  if (!x_dev || !idx_dev || !dist_dev || !ws || (narrow && !idx32_dev)) return dc.last_rc;
  cudaStream_t st = dc.compute_stream();
  if (!st) return VVB_ECUDA;
  const int64_t launches0 = g_launches.load();
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rows = nq;
    stats->algo_used = use;
  }
  rc = dc.copier.copy(x_dev, x_host, static_cast<size_t>(n) * d * sizeof(float), cudaMemcpyHostToDevice, nullptr);
  if (rc) return rc;
  // The D2H copy of chunk i runs on a helper thread (multi-threaded pinned staging) while the kernels of
  // chunk i+1 execute; one copy at a time.
  int dev = 0;
  VVB_CUDA(cudaGetDevice(&dev));
  std::thread bg;
  int bg_rc = VVB_OK;
  std::vector<cudaEvent_t> events;
  bool joined = false;
  auto flush = [&](int64_t c0, int64_t rows, cudaEvent_t ev) -> int {
    if (bg.joinable()) bg.join();
    if (bg_rc) return bg_rc;
    if (!joined) {
      pf.join();
      joined = true;
    }
    bg = std::thread([&dc, &bg_rc, dev, c0, rows, ev, k, idx_out_host, dist_out_host, idx_dev, idx32_dev, dist_dev]() {
      cudaSetDevice(dev);
      int r = idx32_dev ? dc.copier.copy(idx_out_host + c0 * (k + 1), idx32_dev + c0 * (k + 1),
                                         static_cast<size_t>(rows) * (k + 1) * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, ev, CONV_WIDEN_I32)
                        : dc.copier.copy(idx_out_host + c0 * (k + 1), idx_dev + c0 * (k + 1),
                                         static_cast<size_t>(rows) * (k + 1) * sizeof(int64_t),
                                         cudaMemcpyDeviceToHost, ev);
      if (!r)
        r = dc.copier.copy(dist_out_host + c0 * (k + 1), dist_dev + c0 * (k + 1),
                           static_cast<size_t>(rows) * (k + 1) * sizeof(float), cudaMemcpyDeviceToHost, ev);
      bg_rc = r;
    });
    return VVB_OK;
  };
  int64_t pend_c0 = -1, pend_rows = 0;
  cudaEvent_t pend_ev = nullptr;
  rc = knn_run(x_dev, n, d, k, q0, nq, use, idx_dev, dist_dev, ws, ws_bytes, st, stats, /*host_path=*/true,
               [&](int64_t c0, int64_t rows) -> int {
                 if (idx32_dev && rows > 0) {  // right behind the chunk's kernels, ahead of the next chunk's
                   const int64_t cnt = rows * (k + 1);
                   narrow_idx_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(
                       idx_dev + c0 * (k + 1), idx32_dev + c0 * (k + 1), cnt);
                   VVB_CHECK_LAUNCH();
                 }
                 cudaEvent_t e = nullptr;
                 VVB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                 events.push_back(e);
                 VVB_CUDA(cudaEventRecord(e, st));
                 int r2 = VVB_OK;
                 if (pend_c0 >= 0) r2 = flush(pend_c0, pend_rows, pend_ev);
                 pend_ev = e;
                 pend_c0 = c0;
                 pend_rows = rows;
                 return r2;
               });
  if (!rc && pend_c0 >= 0) rc = flush(pend_c0, pend_rows, pend_ev);
  if (bg.joinable()) bg.join();
  if (!rc) rc = bg_rc;
  if (!joined) pf.join();
  cudaError_t e1 = cudaStreamSynchronize(st);
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (rc) {
    if (rc == bg_rc) set_error("make_knn: staged device-to-host copy failed");
    return rc;
  }
  if (e1 != cudaSuccess) {
    set_error("make_knn: %s", cudaGetErrorString(e1));
    return VVB_ECUDA;
  }
  if (stats) stats->kernel_launches = g_launches.load() - launches0;
  return VVB_OK;
}

// The two O(n) scans of ensure_valid_knn (knn.py:58-61: self in column 0, zero distance to self), strided
// over a row-major matrix: one cache line per row and array, split over a few threads.
int vvb_knn_validate_host(const int64_t *idx_host, int64_t ld_idx, const void *dist_host, int dist_elem_size,
                          int64_t ld_dist, int64_t n, int *self_first_out, int *dist0_zero_out) {
  VVB_REQUIRE(idx_host && dist_host && self_first_out && dist0_zero_out, "knn validate: null pointer argument");
  VVB_REQUIRE(n >= 0 && ld_idx >= 1 && ld_dist >= 1 && (dist_elem_size == 4 || dist_elem_size == 8),
              "knn validate: bad shape arguments");
  const int nt = n >= (1 << 16) ? 8 : 1;
  std::atomic<int> self_ok{1}, zero_ok{1};
  auto work = [&](int t) {
    const int64_t r0 = n * t / nt, r1 = n * (t + 1) / nt;
    bool s_ok = true, z_ok = true;
    for (int64_t i = r0; i < r1; ++i) s_ok &= (idx_host[i * ld_idx] == i);
    if (dist_elem_size == 4) {
      const float *dp = static_cast<const float *>(dist_host);
      for (int64_t i = r0; i < r1; ++i) z_ok &= (dp[i * ld_dist] == 0.0f);
    } else {
      const double *dp = static_cast<const double *>(dist_host);
      for (int64_t i = r0; i < r1; ++i) z_ok &= (dp[i * ld_dist] == 0.0);
    }
    if (!s_ok) self_ok = 0;
    if (!z_ok) zero_ok = 0;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  *self_first_out = self_ok.load();
  *dist0_zero_out = zero_ok.load();
  return VVB_OK;
}

int vvb_host_alloc(size_t bytes, void **ptr_out) {
  VVB_REQUIRE(ptr_out && bytes > 0, "host alloc: bad arguments");
  *ptr_out = nullptr;
  cudaError_t e = cudaHostAlloc(ptr_out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    *ptr_out = nullptr;
    set_error("cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VVB_ENOMEM : VVB_ECUDA;
  }
  return VVB_OK;
}

int vvb_host_free(void *ptr) {
  if (!ptr) return VVB_OK;
  VVB_CUDA(cudaFreeHost(ptr));
  return VVB_OK;
}

int vvb_release_cached_memory(void) {
  DeviceCache &dc = DeviceCache::get();
  std::lock_guard<std::mutex> lock(dc.mu);
  dc.release();
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
  const size_t ib = static_cast<size_t>(n) * ld_idx * sizeof(int64_t);
  HostPrefault pf;
  pf.add(out_host, xb);
  pf.start();
  DeviceCache &dc = DeviceCache::get();
  std::lock_guard<std::mutex> lock(dc.mu);
  dc.begin_request();
  void *x = dc.slot(0, xb), *out = dc.slot(4, xb), *idx = dc.slot(1, ib), *ws = dc.slot(3, ws_bytes);
  if (!x || !out || !idx || !ws) return dc.last_rc;
  cudaStream_t st = dc.compute_stream();
  if (!st) return VVB_ECUDA;
  // the whole index matrix goes over contiguously (a pitched copy of the first k columns moves fewer
  // bytes but runs far slower from pageable memory); the kernel strides by ld_idx
  // A page-locked neighbour matrix (one this package returned) goes over as one DMA on the copy stream while
  // the CPU threads stage the pageable `x`.
  const bool idx_pinned = host_range_is_pinned(idx_host, ib);
  cudaEvent_t idx_ev = nullptr;
  if (idx_pinned) {
    cudaStream_t cs = dc.copy_stream();
    if (!cs) return VVB_ECUDA;
    VVB_CUDA(cudaEventCreateWithFlags(&idx_ev, cudaEventDisableTiming));
    VVB_CUDA(cudaMemcpyAsync(idx, idx_host, ib, cudaMemcpyHostToDevice, cs));
    VVB_CUDA(cudaEventRecord(idx_ev, cs));
  }
  struct EvGuard {
    cudaEvent_t &e;
    ~EvGuard() {
      if (e) {
        cudaEventSynchronize(e);  // never leave with the DMA still reading the caller's array
        cudaEventDestroy(e);
      }
    }
  } idx_ev_guard{idx_ev};
  rc = dc.copier.copy(x, x_host, xb, cudaMemcpyHostToDevice, nullptr);
  if (rc) return rc;
  // Opt-in (VVB_SMOOTH_PIPELINE=1): measured neutral on the bench host (1M x 50: 42 ms against 40 ms for the
  // plain sequence) -- the staging memcpys of both directions share the host's memory bandwidth (~20 GB/s
  // net), PCIe (54 GB/s pinned, each way) is not the limit.  Pays off on hosts with faster memory.
  const char *pipe_env = getenv("VVB_SMOOTH_PIPELINE");
  const bool pipeline = n_iter == 1 && ib >= (32u << 20) && pipe_env && atoi(pipe_env) == 1 && !idx_pinned;
  if (pipeline) {
    // One sweep: rows are final as soon as they are written and (Gauss-Seidel) only read rows below them, so
    // the sweep runs chunk by chunk while a helper thread uploads the neighbour rows of the next chunk and
    // this thread downloads the finished ones -- PCIe carries both directions at once.
    VVB_REQUIRE(k <= n - 1, "`k` must be between 1 and (n-1)");
    VVB_REQUIRE(coef >= 0.0 && coef <= 1.0, "`coef` must be more than 0. and at most 1.");
    int nchunks = static_cast<int>(ib / (48u << 20));
    if (nchunks > smooth_max_chunks()) nchunks = smooth_max_chunks();
    if (nchunks < 2) nchunks = 2;
    const int64_t per = (n + nchunks - 1) / nchunks;
    int dev = 0;
    VVB_CUDA(cudaGetDevice(&dev));
    rc = smooth_begin(ws, elem_size, n, d, 1, st);
    if (rc) return rc;
    std::vector<cudaEvent_t> ev(nchunks, nullptr);
    for (auto &e : ev) VVB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    std::atomic<int> ready{0}, prc{VVB_OK};
    std::string perr;
    std::thread producer([&]() {
      cudaSetDevice(dev);
      for (int c = 0; c < nchunks; ++c) {
        const int64_t r0 = std::min<int64_t>(n, per * c), r1 = std::min<int64_t>(n, r0 + per);
        int r = dc.copier2.copy(static_cast<char *>(idx) + static_cast<size_t>(r0) * ld_idx * sizeof(int64_t),
                                reinterpret_cast<const char *>(idx_host) + static_cast<size_t>(r0) * ld_idx * sizeof(int64_t),
                                static_cast<size_t>(r1 - r0) * ld_idx * sizeof(int64_t), cudaMemcpyHostToDevice, nullptr);
        if (!r)
          r = launch_smooth_rows(x, elem_size, n, d, static_cast<const int64_t *>(idx), ld_idx, k, coef, mode, out, ws,
                                 ws_bytes, c, r0, r1, st);
        if (!r && cudaEventRecord(ev[c], st) != cudaSuccess) r = VVB_ECUDA;
        if (r) {
          perr = vvb_last_error();  // the message is per thread: hand it to the caller's
          prc = r;
          ready = nchunks;
          return;
        }
        ready = c + 1;
      }
    });
    pf.join();
    int crc = VVB_OK;
    for (int c = 0; c < nchunks && !crc; ++c) {
      while (ready.load() <= c) std::this_thread::yield();
      if (prc.load()) break;
      const int64_t r0 = std::min<int64_t>(n, per * c), r1 = std::min<int64_t>(n, r0 + per);
      crc = dc.copier.copy(static_cast<char *>(out_host) + static_cast<size_t>(r0) * d * elem_size,
                           static_cast<const char *>(out) + static_cast<size_t>(r0) * d * elem_size,
                           static_cast<size_t>(r1 - r0) * d * elem_size, cudaMemcpyDeviceToHost, ev[c]);
    }
    producer.join();
    int bad = 0;
    cudaError_t ce = cudaMemcpyAsync(&bad, ws, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    for (auto &e : ev) cudaEventDestroy(e);
    if (prc.load()) {
      set_error("%s", perr.c_str());
      return prc.load();
    }
    if (crc) return crc;
    VVB_CUDA(ce);
    VVB_REQUIRE(bad == 0, "Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed");
    return VVB_OK;
  }
  if (idx_pinned) {
    VVB_CUDA(cudaStreamWaitEvent(st, idx_ev, 0));
  } else if (n <= INT32_MAX) {
    void *idx32 = dc.slot(5, ib / 2);
    if (!idx32) return dc.last_rc;
    rc = dc.copier.copy(idx32, idx_host, ib / 2, cudaMemcpyHostToDevice, nullptr, CONV_NARROW_I64);
    if (rc) return rc;
    const int64_t cnt = n * ld_idx;
    widen_idx_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(static_cast<const int32_t *>(idx32),
                                                                              static_cast<int64_t *>(idx), cnt);
    VVB_CHECK_LAUNCH();
  } else {
    rc = dc.copier.copy(idx, idx_host, ib, cudaMemcpyHostToDevice, nullptr);
    if (rc) return rc;
  }
  rc = vvb_smooth_device(x, elem_size, n, d, static_cast<const int64_t *>(idx), ld_idx, k, coef, n_iter, mode, out, ws,
                         ws_bytes, st);
  if (rc) return rc;
  int bad = 0;
  VVB_CUDA(cudaMemcpyAsync(&bad, ws, sizeof(int), cudaMemcpyDeviceToHost, st));
  cudaEvent_t done = nullptr;
  VVB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  VVB_CUDA(cudaEventRecord(done, st));
  pf.join();
  rc = dc.copier.copy(out_host, out, xb, cudaMemcpyDeviceToHost, done);
  cudaEventDestroy(done);
  if (rc) return rc;
  VVB_CUDA(cudaStreamSynchronize(st));
  VVB_REQUIRE(bad == 0, "Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed");
  return VVB_OK;
}

int vvb_smooth_cols_device(const void *x_dev, int elem_size, int64_t n, int d, int col0, int dcols,
                           const int64_t *idx_dev, int64_t ld_idx, int k, double coef, int n_iter, int mode,
                           void *out_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
  VVB_REQUIRE(x_dev && idx_dev && out_dev && workspace_dev, "smooth: null pointer argument");
  VVB_REQUIRE(k >= 1 && k <= n - 1, "`k` must be between 1 and (n-1)");
  VVB_REQUIRE(coef >= 0.0 && coef <= 1.0, "`coef` must be more than 0. and at most 1.");
  VVB_REQUIRE(n_iter >= 1 && n_iter <= 10000, "`n_iter` must be at least 1 and at most 10000");
  return launch_smooth_cols(x_dev, elem_size, n, d, col0, dcols, idx_dev, ld_idx, k, coef, n_iter, mode, out_dev,
                            workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

// Rows [row_lo, row_hi) of ONE Jacobi sweep: every row reads only `x`, so row ranges are independent and can
// live on different GPUs.  idx_rows / out_rows hold just those rows.
int vvb_smooth_rows_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_rows_dev,
                           int64_t ld_idx, int k, double coef, int mode, int64_t row_lo, int64_t row_hi,
                           void *out_rows_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
  VVB_REQUIRE(x_dev && idx_rows_dev && out_rows_dev && workspace_dev, "smooth rows: null pointer argument");
  VVB_REQUIRE(k >= 1 && k <= n - 1, "`k` must be between 1 and (n-1)");
  VVB_REQUIRE(coef >= 0.0 && coef <= 1.0, "`coef` must be more than 0. and at most 1.");
  VVB_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= n, "smooth rows: bad row range");
  VVB_REQUIRE(elem_size == 4 || elem_size == 8, "smooth: elem_size must be 4 or 8");
  if (mode != VVB_SMOOTH_JACOBI) {
    set_error("smooth rows: only VVB_SMOOTH_JACOBI sweeps split by rows (a Gauss-Seidel row depends on the rows below it)");
    return VVB_ENOTSUP;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = smooth_begin(workspace_dev, elem_size, n, d, 1, st);
  if (rc) return rc;
  // the kernel addresses rows globally: hand it the bases the full matrices would have
  const int64_t *idx_base = idx_rows_dev - row_lo * ld_idx;
  char *out_base = static_cast<char *>(out_rows_dev) - static_cast<size_t>(row_lo) * d * elem_size;
  return launch_smooth_rows(x_dev, elem_size, n, d, idx_base, ld_idx, k, coef, mode, out_base, workspace_dev,
                            workspace_bytes, 0, row_lo, row_hi, st);
}

// -------------------------------------------------------------------- smoother over peer memory (several GPUs)
int vvb_peer_alloc(size_t bytes, void **ptr_out, unsigned char *handle_out) {
  VVB_REQUIRE(ptr_out && handle_out && bytes > 0, "peer alloc: bad arguments");
  *ptr_out = nullptr;
  cudaError_t e = cudaMalloc(ptr_out, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VVB_ENOMEM : VVB_ECUDA;
  }
  VVB_CUDA(cudaMemset(*ptr_out, 0, bytes));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr_out);
  if (e != cudaSuccess) {
    cudaFree(*ptr_out);
    *ptr_out = nullptr;
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return VVB_ECUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  return VVB_OK;
}

int vvb_peer_open(const unsigned char *handle, void **ptr_out) {
  VVB_REQUIRE(handle && ptr_out, "peer open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  *ptr_out = nullptr;
  VVB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return VVB_OK;
}

int vvb_peer_close(void *ptr) {
  if (ptr) VVB_CUDA(cudaIpcCloseMemHandle(ptr));
  return VVB_OK;
}

int vvb_peer_free(void *ptr) {
  if (ptr) VVB_CUDA(cudaFree(ptr));
  return VVB_OK;
}

size_t vvb_smooth_peer_workspace_bytes(void) { return smooth_peer_workspace_bytes(); }

int vvb_smooth_peer_device(const void *x_dev, int elem_size, int64_t n, int d, const int64_t *idx_dev, int64_t ld_idx,
                           int k, double coef, int epoch, int rank, int world, void *const *ll_ptrs,
                           void *workspace_dev, void *stream) {
  VVB_REQUIRE(x_dev && idx_dev && ll_ptrs && workspace_dev, "smooth (peer): null pointer argument");
  VVB_REQUIRE(k >= 1 && k <= n - 1, "`k` must be between 1 and (n-1)");
  VVB_REQUIRE(coef >= 0.0 && coef <= 1.0, "`coef` must be more than 0. and at most 1.");
  return launch_smooth_peer(x_dev, elem_size, n, d, idx_dev, ld_idx, k, coef, epoch, rank, world, ll_ptrs, workspace_dev,
                            static_cast<cudaStream_t>(stream));
}

int vvb_smooth_peer_unpack_device(const void *ll_local_dev, int elem_size, int64_t count, int epoch, void *out_dev,
                                  void *workspace_dev, void *stream) {
  VVB_REQUIRE(ll_local_dev && out_dev && workspace_dev && count >= 0, "smooth (peer) unpack: bad arguments");
  return launch_smooth_peer_unpack(ll_local_dev, elem_size, count, epoch, out_dev, workspace_dev,
                                   static_cast<cudaStream_t>(stream));
}

// -------------------------------------------------------------------- ring merge
int vvb_knn_merge_init_device(int64_t *run_idx_dev, float *run_d2_dev, int64_t nq, int k, int64_t self0, void *stream) {
  VVB_REQUIRE(run_idx_dev && run_d2_dev, "knn merge init: null pointer argument");
  return launch_knn_merge_init(run_idx_dev, run_d2_dev, nq, k, self0, static_cast<cudaStream_t>(stream));
}

int vvb_knn_merge_device(const float *xc_dev, int64_t nc, int d, int64_t q0, int64_t nq, const int64_t *new_idx_dev,
                         int kn, int64_t keep_lo, int64_t keep_hi, int64_t global_off, int64_t *run_idx_dev,
                         float *run_d2_dev, int k, void *stream) {
  VVB_REQUIRE(xc_dev && new_idx_dev && run_idx_dev && run_d2_dev, "knn merge: null pointer argument");
  return launch_knn_merge(xc_dev, nc, d, q0, nq, new_idx_dev, kn, keep_lo, keep_hi, global_off, run_idx_dev,
                          run_d2_dev, k, static_cast<cudaStream_t>(stream));
}

int vvb_knn_merge_finish_device(const float *run_d2_dev, int64_t count, float *dist_out_dev, void *stream) {
  VVB_REQUIRE(run_d2_dev && dist_out_dev, "knn merge finish: null pointer argument");
  return launch_knn_merge_finish(run_d2_dev, count, dist_out_dev, static_cast<cudaStream_t>(stream));
}

// -------------------------------------------------------------------- fused training step
int vvb_vae_step_device(const float *const *w_ptrs, const float *const *b_ptrs, const int *in_dims, const int *out_dims,
                        int n_enc, int n_dec, const float *x_dev, const float *eps_dev, int B, int D, int L,
                        float lam_recon, float lam_kldiv, float lam_mds, float *grad_partials_dev,
                        float *loss_partials_dev, float *grad_flat_dev, float *term_sums_dev, float *loss_out_dev,
                        size_t *smem_bytes, int *n_params, int *n_cta, void *stream) {
  VVB_REQUIRE(in_dims && out_dims, "fused step: null dimension arrays");
  if (grad_partials_dev)
    VVB_REQUIRE(w_ptrs && b_ptrs && x_dev && eps_dev && loss_partials_dev && grad_flat_dev && loss_out_dev,
                "fused step: null pointer argument");
  return launch_vae_step(w_ptrs, b_ptrs, in_dims, out_dims, n_enc, n_dec, x_dev, eps_dev, B, D, L, lam_recon, lam_kldiv,
                         lam_mds, grad_partials_dev, loss_partials_dev, grad_flat_dev, term_sums_dev, loss_out_dev,
                         smem_bytes, n_params, n_cta, static_cast<cudaStream_t>(stream));
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
