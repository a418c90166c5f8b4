// libkemr.so -- C ABI (include/kemr.h) over the sm_100a kernels.  Host logic only: argument
// checks, workspace carving, kernel selection and launches.  Nothing here synchronises the
// device except the kemr_index_*_host calls.
#include "../../include/kemr.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <algorithm>

#include "common.cuh"
#include "scan_warp.cuh"
#include "select.cuh"
#include "matrix_ops.cuh"
#include "scan_mma.cuh"
#include "scan_stream.cuh"
#include "store.cuh"

#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <string_view>
#include <unordered_map>
#include <mutex>

using namespace kemr;

// ----------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(KEMR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                  __FILE__, __LINE__);                                                      \
  } while (0)
#define LAUNCH_CHECK(what)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return fail(KEMR_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e__)); \
  } while (0)

extern "C" const char* kemr_last_error(void) { return g_last_error.c_str(); }

// measurement hook: when set, the event is recorded right after the scan kernel of the next
// kemr_scan_topk / kemr_rank_count calls of this thread (lets bench.py time the scan kernel alone)
static thread_local cudaEvent_t g_scan_done_event = nullptr;
extern "C" int kemr_set_scan_done_event(void* cuda_event) {
  g_scan_done_event = reinterpret_cast<cudaEvent_t>(cuda_event);
  return KEMR_OK;
}
// measurement hook: device int64[3] that the fused small-batch kernel stamps with %globaltimer (first CTA start,
// last scan arrival, selection done) on the following kemr_scan_topk calls of this thread
static thread_local long long* g_phase_stamps = nullptr;
extern "C" int kemr_set_phase_stamps(void* device_int64x3) {
  g_phase_stamps = reinterpret_cast<long long*>(device_int64x3);
  return KEMR_OK;
}
extern "C" int kemr_abi_version(void) { return KEMR_ABI_VERSION; }

// Arrival counters of the fused small-batch kernel ("last CTA runs the selection"): a per-device pool, zeroed once;
// every launch takes the next slice and its last CTA resets what it used, so launches on different streams never
// share a counter (until 64 Ki groups later, long after the earlier launch has drained).
static int stream_counters(int groups, unsigned int** out) {
  constexpr int kPool = 1 << 16;
  struct Pool { unsigned int* p = nullptr; int dev = -1; unsigned int cursor = 0; };
  static Pool pools[64];
  static std::mutex mu;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || groups > kPool) return fail(KEMR_ERR_UNSUPPORTED, "stream_counters: device %d / %d groups", dev, groups);
  std::lock_guard<std::mutex> lock(mu);
  Pool& pl = pools[dev];
  if (!pl.p) {
    CUDA_TRY(cudaMalloc(&pl.p, (size_t)kPool * 4));
    CUDA_TRY(cudaMemset(pl.p, 0, (size_t)kPool * 4));
  }
  if (pl.cursor + (unsigned)groups > (unsigned)kPool) pl.cursor = 0;
  *out = pl.p + pl.cursor;
  pl.cursor += (unsigned)groups;
  return KEMR_OK;
}

struct DevInfo { int ok = 0, dev = -1, sms = 0, major = 0, minor = 0, quads = 0; };
static int dev_info(DevInfo* out) {
  static thread_local DevInfo cache;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!cache.ok || cache.dev != dev) {
    cache.dev = dev;
    CUDA_TRY(cudaDeviceGetAttribute(&cache.sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cache.major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cache.minor, cudaDevAttrComputeCapabilityMinor, dev));
    cache.quads = cache.major == 10 ? mma_max_quads() : 0;
    cache.ok = 1;
  }
  *out = cache;
  return KEMR_OK;
}

extern "C" int kemr_device_info(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05) {
  DevInfo d;
  int rc = dev_info(&d);
  if (rc) return rc;
  if (sm_count) *sm_count = d.sms;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  if (has_tcgen05) *has_tcgen05 = (d.major == 10 && mma_built()) ? 1 : 0;
  return KEMR_OK;
}

static inline cudaStream_t S(kemr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// The tcgen05 scan needs the SM's whole shared-memory carve-out (227 KB).  The small kernels around it (quantise,
// selection, merge) ask for the same carve-out, so that an SM does not have to re-partition L1 / shared memory
// between the kernels of one search step.  Once per function and thread.
template <class F>
static inline void prefer_max_smem(F* func) {
  // per kernel (instantiations of one template share F, so the flag cannot be a static of this function template)
  static thread_local std::vector<const void*> done;
  const void* key = reinterpret_cast<const void*>(func);
  for (const void* p : done) if (p == key) return;
  cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();
  done.push_back(key);
}
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------- simple kernels
extern "C" int kemr_quantize_rows(const float* src, uint16_t* dst, int64_t rows, int D, int normalize,
                                  kemr_stream_t stream) {
  if (!src || !dst || rows < 0 || D <= 0) return fail(KEMR_ERR_ARG, "quantize_rows: bad argument");
  if (rows == 0) return KEMR_OK;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, 148 * 16);
  if (!normalize && D % 128 == 0 && D <= 1024 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    switch (D / 128) {
#define KEMR_QV(n) case n: prefer_max_smem(quantize_rows_vec_kernel<n>); quantize_rows_vec_kernel<n><<<(unsigned)blocks, threads, 0, S(stream)>>>(src, dst, rows); break;
      KEMR_QV(1) KEMR_QV(2) KEMR_QV(3) KEMR_QV(4) KEMR_QV(5) KEMR_QV(6) KEMR_QV(7) KEMR_QV(8)
#undef KEMR_QV
    }
    LAUNCH_CHECK("quantize_rows_vec_kernel");
    return KEMR_OK;
  }
  prefer_max_smem(quantize_rows_kernel);
  quantize_rows_kernel<<<(unsigned)blocks, threads, 0, S(stream)>>>(src, dst, rows, D, normalize);
  LAUNCH_CHECK("quantize_rows_kernel");
  return KEMR_OK;
}

extern "C" int kemr_row_norm_max(const uint16_t* x, int64_t rows, int D, float* out_max, kemr_stream_t stream) {
  if (!x || !out_max || rows < 0 || D <= 0 || D % 8) return fail(KEMR_ERR_ARG, "row_norm_max: bad argument");
  CUDA_TRY(cudaMemsetAsync(out_max, 0, 4, S(stream)));
  if (rows == 0) return KEMR_OK;
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, 148 * 16);
  row_norm_max_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(x, rows, D, out_max);
  LAUNCH_CHECK("row_norm_max_kernel");
  return KEMR_OK;
}

extern "C" int kemr_synth_rows(uint16_t* dst, int64_t rows, int D, uint64_t seed, int64_t row_base,
                               kemr_stream_t stream) {
  if (!dst || rows < 0 || D <= 0) return fail(KEMR_ERR_ARG, "synth_rows: bad argument");
  if (rows == 0) return KEMR_OK;
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, 148 * 16);
  synth_rows_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(dst, rows, D, seed, row_base);
  LAUNCH_CHECK("synth_rows_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- scan planning
struct ScanPlan {
  int path;        // KEMR_PATH_WARP / KEMR_PATH_MMA
  int QB, CH;      // warp path
  int P;           // parts (lists per query)
  int Kp;          // entries per part list
  int groups;      // warp path: query groups (grid.y)
  MmaPlan mma;     // mma path
};

static int check_common(const void* q, int Q, const void* ga, int64_t M, int D) {
  if (!q || !ga) return fail(KEMR_ERR_ARG, "null embedding pointer");
  if (Q <= 0 || M <= 0) return fail(KEMR_ERR_ARG, "Q and M must be positive (Q=%d, M=%lld)", Q, (long long)M);
  if (M >= (1ll << 31)) return fail(KEMR_ERR_ARG, "a shard holds at most 2^31-1 rows");
  if (D <= 0 || D % 8 != 0 || D > kMaxD) return fail(KEMR_ERR_ARG, "D must be a multiple of 8, <= %d (D=%d)", kMaxD, D);
  if (((uintptr_t)q | (uintptr_t)ga) & 15) return fail(KEMR_ERR_ARG, "embedding pointers must be 16-byte aligned");
  return KEMR_OK;
}

static int make_plan(int Q, int64_t M, int D, int G, int K, int mode, int path, bool equal_weights, const DevInfo& dv, ScanPlan* pl) {
  pl->path = path;
  if (path == KEMR_PATH_AUTO) {
    static const char* force = getenv("KEMR_FORCE_PATH");     // experiments: 1 = warp-dot, 2 = tcgen05
    if (force && (force[0] == '1' || force[0] == '2')) path = force[0] - '0';
    pl->path = path;
  }
  if (path == KEMR_PATH_AUTO) {
    // tensor-core kernel when there is a batch to amortise its prologue and the shape can be planned
    // (enough gallery tiles for the candidate lists k_sel needs); warp-dot for the latency-bound tiny
    // batches and as the general fallback
    pl->path = KEMR_PATH_WARP;
    if (dv.major == 10 && Q >= 3 && mma_supported(D, K) && mma_make_plan(Q, M, D, G, K, mode, dv.sms, dv.quads, equal_weights, &pl->mma) == 0)
      pl->path = KEMR_PATH_MMA;
  }
  if (pl->path == KEMR_PATH_MMA) {
    if (dv.major != 10) return fail(KEMR_ERR_UNSUPPORTED, "tcgen05 path needs compute capability 10.x (have %d.%d)", dv.major, dv.minor);
    if (!mma_supported(D, K)) return fail(KEMR_ERR_UNSUPPORTED, "tcgen05 path: unsupported D=%d / k_sel=%d", D, K);
    int rc = mma_make_plan(Q, M, D, G, K, mode, dv.sms, dv.quads, equal_weights, &pl->mma);
    if (rc) return fail(KEMR_ERR_UNSUPPORTED, "tcgen05 path: cannot plan this shape (k_sel=%d needs more gallery tiles than M=%lld offers)", K, (long long)M);
    pl->P = pl->mma.parts;
    pl->Kp = pl->mma.K;
    return KEMR_OK;
  }
  pl->CH = (D + 255) / 256;
  if (mode == kModeTopk) {
    // fused streaming search (scan_stream.cuh): one 16-warp CTA per SM and query group of 1, 2 or 4 queries
    pl->QB = Q == 1 ? 1 : ((Q == 2 || pl->CH >= 3) ? 2 : 4);      // four queries x 768-d do not fit the registers of 2 CTAs / SM
    pl->groups = (Q + pl->QB - 1) / pl->QB;
    pl->P = (int)std::min<int64_t>(dv.sms > 0 ? dv.sms : 1, std::max<int64_t>(1, (M + 31) / 32));
    pl->Kp = K;
    return KEMR_OK;
  }
  pl->QB = Q == 1 ? 1 : 2;
  pl->groups = (Q + pl->QB - 1) / pl->QB;
  int want = 2 * dv.sms;
  int P = std::max(1, (want + pl->groups - 1) / pl->groups);
  P = (int)std::min<int64_t>(P, std::max<int64_t>(1, (M + 15) / 16));
  pl->P = P;
  pl->Kp = K;
  return KEMR_OK;
}

extern "C" int kemr_scan_plan(int Q, int64_t M, int D, int galleries, int k_sel, int equal_weights, int* path, int* parts) {
  if (Q <= 0 || M <= 0 || D <= 0 || D % 8 || D > kMaxD || galleries < 1 || galleries > 2 || k_sel < 1 || k_sel > kMaxKSel)
    return fail(KEMR_ERR_ARG, "scan_plan: bad argument");
  DevInfo dv;
  int rc = dev_info(&dv);
  if (rc) return rc;
  ScanPlan pl;
  if ((rc = make_plan(Q, M, D, galleries, k_sel, kModeTopk, KEMR_PATH_AUTO, equal_weights != 0, dv, &pl))) return rc;
  if (path) *path = pl.path;
  if (parts) *parts = pl.P;
  return KEMR_OK;
}

static size_t parts_bytes(int P, int Q, int K) { return align_up((size_t)P * Q * K * sizeof(uint64_t)); }

extern "C" size_t kemr_workspace_bytes(int Q, int64_t M, int D, int k_sel, int64_t max_hits_per_query) {
  DevInfo dv;
  int sms = 148;
  if (dev_info(&dv) == KEMR_OK && dv.sms > 0) sms = dv.sms;
  const int P = 2 * sms + 8;                    // upper bound of any plan's part count
  const int K = std::max(1, std::min(k_sel, kMaxKSel));
  const int Qp = (Q + 511) / 512 * 512;       // query rows padded to a cluster's block (two CTA pairs)
  size_t topk = parts_bytes(P, Qp, K);
  size_t count = align_up((size_t)Q * 8) + align_up((size_t)Q * 8) + align_up((size_t)P * Qp * 4) + 256 +
                 align_up(((size_t)1 << 20) * 8 + (size_t)Q * 256 * 8);
  (void)M; (void)D;
  // + per-query weights as fp32, + canonical scores of the KG hits (fused streaming search)
  const size_t hits = align_up((size_t)std::max<int64_t>(0, std::min<int64_t>(max_hits_per_query, 4096)) * (size_t)Q * 8);
  return std::max(topk, count) + 2 * align_up((size_t)Q * 4) + hits + align_up((size_t)Qp * 4) + 4096;
}

template <int QB, int CH>
static void launch_warp(const ScanArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  scan_warp_kernel<QB, CH><<<grid, kWarpScanThreads, smem, st>>>(a);
}
static int launch_warp_scan(const ScanArgs& a, const ScanPlan& pl, cudaStream_t st) {
  dim3 grid(pl.P, pl.groups);
  const size_t smem = 0;
  if (pl.groups > 65535) return fail(KEMR_ERR_UNSUPPORTED, "warp path: too many query groups (%d)", pl.groups);
#define KEMR_CASE(qb, ch) if (pl.QB == qb && pl.CH == ch) { launch_warp<qb, ch>(a, grid, smem, st); }
  KEMR_CASE(1, 1) KEMR_CASE(1, 2) KEMR_CASE(1, 3) KEMR_CASE(1, 4)
  KEMR_CASE(2, 1) KEMR_CASE(2, 2) KEMR_CASE(2, 3) KEMR_CASE(2, 4)
#undef KEMR_CASE
  LAUNCH_CHECK("scan_warp_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- scan + top-k
// per-query weights for the fp32 scan kernels: binary64 [Q] -> fp32 [Q]
__global__ void weights_to_f32_kernel(const double* __restrict__ wa, const double* __restrict__ wb, int Q,
                                      float* __restrict__ oa, float* __restrict__ ob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) { oa[i] = (float)wa[i]; ob[i] = (float)wb[i]; }
}

struct kemr_peer;
static int arm_peer_push(SelectArgs* s, int Q, int k);

static int scan_topk_impl(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                          int64_t M, int D, double w_a, double w_b, const double* wq_a, const double* wq_b, double alpha,
                          const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                          int64_t max_hits_per_query, int k, int k_sel, double eps, int64_t idx_base,
                          double* out_score64, float* out_score32, int64_t* out_idx, int32_t* out_flags,
                          void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream,
                          const float* q_f32 = nullptr, int q_normalize = 0) {
  // q_f32 (internal, host-buffer search of small batches): the queries are still fp32 in device memory and the fused
  // streaming kernel rounds them itself, storing the bf16 rows to `q`; only valid when that kernel is the one chosen
  int rc = check_common(q, Q, gal_a, M, D);
  if (rc) return rc;
  if ((wq_a != nullptr) != (wq_b != nullptr)) return fail(KEMR_ERR_ARG, "per-query weights need both arrays");
  if (wq_a && !gal_b) return fail(KEMR_ERR_ARG, "per-query weights need two galleries");
  if (k <= 0 || k_sel < k || k_sel > kMaxKSel) return fail(KEMR_ERR_ARG, "need 0 < k <= k_sel <= %d (k=%d, k_sel=%d)", kMaxKSel, k, k_sel);
  if (!(alpha > 0.0)) return fail(KEMR_ERR_ARG, "alpha must be > 0 for the sparse KG-boost path (alpha=%g)", alpha);
  if (!out_score64 || !out_idx || !out_flags) return fail(KEMR_ERR_ARG, "null output pointer");
  if (hit_rowptr && (!hit_col || !hit_bonus)) return fail(KEMR_ERR_ARG, "hit CSR needs col and bonus arrays");
  if (!hit_rowptr) max_hits_per_query = 0;
  if (max_hits_per_query < 0 || max_hits_per_query > 4096) return fail(KEMR_ERR_ARG, "max_hits_per_query must be in [0, 4096]");
  DevInfo dv;
  if ((rc = dev_info(&dv))) return rc;
  const int G = gal_b ? 2 : 1;
  ScanPlan pl;
  if ((rc = make_plan(Q, M, D, G, k_sel, kModeTopk, path, !wq_a && (float)w_a == (float)w_b, dv, &pl))) return rc;
  const int Qrows = pl.path == KEMR_PATH_MMA ? pl.mma.q_pad : Q;
  const size_t need = parts_bytes(pl.P, Qrows, pl.Kp);
  const size_t need_w = wq_a ? 2 * align_up((size_t)Q * 4) : 0;
  if (!workspace || workspace_bytes < need + need_w) return fail(KEMR_ERR_WORKSPACE, "scan_topk needs %zu workspace bytes, got %zu", need + need_w, workspace_bytes);
  cudaStream_t st = S(stream);
  uint64_t* part_keys = reinterpret_cast<uint64_t*>(workspace);
  float* wq32[2] = {nullptr, nullptr};
  if (wq_a) {
    wq32[0] = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + need);
    wq32[1] = wq32[0] + align_up((size_t)Q * 4) / 4;
    weights_to_f32_kernel<<<(Q + 255) / 256, 256, 0, st>>>(wq_a, wq_b, Q, wq32[0], wq32[1]);
    LAUNCH_CHECK("weights_to_f32_kernel");
  }

  ScanArgs a{};
  a.q = q; a.Q = Q; a.gal[0] = gal_a; a.gal[1] = gal_b; a.G = G; a.M = M; a.D = D;
  a.w[0] = (float)w_a; a.w[1] = (float)w_b; a.wq[0] = wq32[0]; a.wq[1] = wq32[1];
  a.mode = kModeTopk; a.K = pl.Kp; a.part_keys = part_keys;

  SelectArgs s{};
  s.part_keys = part_keys; s.P = pl.P; s.Q = Qrows; s.K = k_sel; s.Kp = pl.Kp;
  s.q = q; s.gal[0] = gal_a; s.gal[1] = gal_b; s.G = G; s.D = D; s.M = M;
  s.w[0] = w_a; s.w[1] = w_b; s.wq[0] = wq_a; s.wq[1] = wq_b; s.alpha = alpha;
  s.hit_rowptr = hit_rowptr; s.hit_col = hit_col; s.hit_bonus = hit_bonus;
  s.k = k; s.eps = eps; s.idx_base = idx_base;
  s.out_score64 = out_score64; s.out_score32 = out_score32; s.out_idx = out_idx; s.out_flags = out_flags;
  s.max_cand = k_sel + (int)max_hits_per_query;
  s.key_slots = (int)select_key_slots(pl.Kp, k_sel);
  if ((rc = arm_peer_push(&s, Q, k))) return rc;
  if (pl.P > kMaxParts) return fail(KEMR_ERR_UNSUPPORTED, "too many part lists per query (%d)", pl.P);
  const size_t smem = select_smem_bytes(pl.P, pl.Kp, k_sel, s.max_cand);
  if (smem > 200 * 1024) return fail(KEMR_ERR_UNSUPPORTED, "select kernel needs %zu bytes of shared memory", smem);
  const int np = (D + 255) / 256;

  if (pl.path != KEMR_PATH_MMA) {
    // small batches: scan + selection in ONE launch (scan_stream.cuh)
    StreamArgs sa{};
    sa.s = a; sa.sel = s;
    const size_t dyn = stream_smem_bytes(pl.QB, smem);
    sa.cand_off = (unsigned)stream_cand_off(smem);
    if (dyn > (size_t)kSmemBudget) return fail(KEMR_ERR_UNSUPPORTED, "stream kernel needs %zu bytes of shared memory", dyn);
    if (pl.groups > 65535) return fail(KEMR_ERR_UNSUPPORTED, "warp path: too many query groups (%d)", pl.groups);
    if ((rc = stream_counters(pl.groups, &sa.done))) return rc;
    sa.stamps = g_phase_stamps;
    sa.q_f32 = q_f32; sa.q_normalize = q_normalize;
    // canonical scores of the KG hits, computed by the CTAs as they finish their scan (scratch behind the part lists;
    // skipped when the caller's workspace has no room for it)
    const size_t hit_bytes = hit_rowptr ? align_up((size_t)Q * (size_t)max_hits_per_query * 8) : 0;
    if (hit_bytes && workspace_bytes >= need + need_w + hit_bytes)
      sa.sel.hit_score = reinterpret_cast<const double*>(reinterpret_cast<unsigned char*>(workspace) + need + need_w);
    if (sa.stamps) CUDA_TRY(cudaMemsetAsync(sa.stamps, 0x7f, 8, st));       // "first start" is an atomicMin
    dim3 grid(pl.P, pl.groups);
#define KEMR_STREAM(QBV, CHV)                                                                                     \
  do {                                                                                                            \
    /* the attribute sticks to the function on a device: set it once per (instantiation, device, size) */         \
    static thread_local int attr_dev = -1;                                                                        \
    static thread_local size_t attr_smem = 0;                                                                     \
    if (dv.dev != attr_dev || dyn > attr_smem) {                                                                  \
      CUDA_TRY(cudaFuncSetAttribute(scan_stream_kernel<QBV, CHV, CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
      attr_dev = dv.dev; attr_smem = dyn;                                                                         \
    }                                                                                                             \
    scan_stream_kernel<QBV, CHV, CHV><<<grid, kStreamThreads, dyn, st>>>(sa);                                     \
  } while (0)
#define KEMR_STREAM_CH(QBV) switch (np) { case 1: KEMR_STREAM(QBV, 1); break; case 2: KEMR_STREAM(QBV, 2); break; \
                                          case 3: KEMR_STREAM(QBV, 3); break; default: KEMR_STREAM(QBV, 4); break; }
    if (pl.QB == 1) { KEMR_STREAM_CH(1) } else if (pl.QB == 2) { KEMR_STREAM_CH(2) } else { KEMR_STREAM_CH(4) }
#undef KEMR_STREAM_CH
#undef KEMR_STREAM
    LAUNCH_CHECK("scan_stream_kernel");
    if (g_scan_done_event) CUDA_TRY(cudaEventRecord(g_scan_done_event, st));
    return KEMR_OK;
  }

  if (q_f32) return fail(KEMR_ERR_ARG, "internal: fp32 queries need the fused streaming kernel");
  if (!pl.mma.all_slots) CUDA_TRY(cudaMemsetAsync(part_keys, 0, need, st));   // unwritten slots must read as empty
  // per-query threshold shared by the lists of a query (scan_mma.cuh): zero = nothing published; behind the part lists
  // and the fp32 weights, when the caller's workspace has the room
  // Measured (profiles/r02_session_l2_stdout.txt): +4.4 % on 4096 queries x 1.25 M rows, top-100 (short lists that
  // restart at virtual part boundaries start warm: 5.78 -> 5.53 ms); no gain where every list of a query lives for the
  // whole scan (C1: 198.4 us without, 200.4 us with; C2 the same to 0.2 us) -- all lists warm up at the same pace, so
  // the neighbours' thresholds are no better than a list's own and the loads are overhead.  So: plans with virtual parts.
  static const char* share_env = getenv("KEMR_THR_SHARE");          // experiments: 0 = never, 1 = always
  const bool share = share_env ? share_env[0] == '1' : pl.mma.vq > 1;
  const size_t thr_bytes = align_up((size_t)Qrows * 4);
  if (share && workspace_bytes >= need + need_w + thr_bytes) {
    a.thr_pub = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(workspace) + need + need_w);
    CUDA_TRY(cudaMemsetAsync(a.thr_pub, 0, (size_t)Qrows * 4, st));
  }
  if ((rc = mma_launch(a, pl.mma, st))) return fail(KEMR_ERR_CUDA, "tcgen05 scan launch failed: %s", mma_last_error());
  if (g_scan_done_event) CUDA_TRY(cudaEventRecord(g_scan_done_event, st));

  // one CTA per query: 4 warps for the common small case, 8 when there are many candidates to re-score
  // Programmatic dependent launch: the selection's CTAs become resident while the scan drains (they block in
  // griddepcontrol.wait), so its launch latency and ramp hide behind the scan's tail.  Not when an event has to be
  // recorded between the two kernels (scan timing hook) -- anything between them in the stream makes it an ordinary launch.
  static const bool no_pdl = getenv("KEMR_NO_PDL") != nullptr;
  const bool pdl = !no_pdl && !g_scan_done_event;
#define KEMR_SEL(NPV, WV)                                                                                         \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      CUDA_TRY(cudaFuncSetAttribute(select_kernel<NPV, WV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    prefer_max_smem(select_kernel<NPV, WV>);                                                                      \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = dim3((unsigned)Q); cfg.blockDim = dim3(WV * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;  \
    cudaLaunchAttribute at[1];                                                                                    \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                                         \
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;                                                                   \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, select_kernel<NPV, WV>, s));                                                \
  } while (0)
#define KEMR_SEL_NP(WV) switch (np) { case 1: KEMR_SEL(1, WV); break; case 2: KEMR_SEL(2, WV); break; \
                                      case 3: KEMR_SEL(3, WV); break; default: KEMR_SEL(4, WV); break; }
  // Many queries (several waves of 4-warp CTAs at four CTAs per SM): two warps per query -- the selection is a chain
  // of dependent phases, so narrower CTAs in twice the number keep more queries in flight.  A/B on one box
  // (profiles/r02_session_ab_select_warps.txt): C1 (4300 queries) 52.7 -> 40.5 us; C2 (1000 queries, under two waves)
  // 26.8 -> 33.0 us, so it stays at four.
  static const int sel_w = getenv("KEMR_SEL_W") ? atoi(getenv("KEMR_SEL_W")) : 0;               // experiments: 2 / 4 / 8
  // Few queries leave the GPU nearly empty: sixteen warps per query while every query gets an SM of its own (one pass
  // over the lists' heads, one round of re-scoring), eight up to two per SM.  Batch 4 ... 128 at 43 000 rows: 57-65 ->
  // 52-54 us per step (profiles/r02_session_x_select16.txt); C2's 1000 queries: 4 warps 26.6 us, 8 warps 37.3, 16 warps 53.
  int W = 8;
  if (Q <= dv.sms) W = 16;
  else if (Q <= 2 * dv.sms || s.max_cand > kSelSmallCand) W = 8;
  else W = Q > 12 * dv.sms ? 2 : 4;
  if (sel_w) W = sel_w;
  if (W == 2) { KEMR_SEL_NP(2) } else if (W == 4) { KEMR_SEL_NP(4) } else if (W == 16) { KEMR_SEL_NP(16) } else { KEMR_SEL_NP(8) }
#undef KEMR_SEL_NP
#undef KEMR_SEL
  LAUNCH_CHECK("select_kernel");
  return KEMR_OK;
}

extern "C" int kemr_scan_topk(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                              int64_t M, int D, double w_a, double w_b, double alpha,
                              const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                              int64_t max_hits_per_query, int k, int k_sel, double eps, int64_t idx_base,
                              double* out_score64, float* out_score32, int64_t* out_idx, int32_t* out_flags,
                              void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream) {
  return scan_topk_impl(q, Q, gal_a, gal_b, M, D, w_a, w_b, nullptr, nullptr, alpha, hit_rowptr, hit_col, hit_bonus,
                        max_hits_per_query, k, k_sel, eps, idx_base, out_score64, out_score32, out_idx, out_flags,
                        workspace, workspace_bytes, path, stream);
}

extern "C" int kemr_scan_topk_gated(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                                    int64_t M, int D, const double* w_a_q, const double* w_b_q, double alpha,
                                    const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                                    int64_t max_hits_per_query, int k, int k_sel, double eps, int64_t idx_base,
                                    double* out_score64, float* out_score32, int64_t* out_idx, int32_t* out_flags,
                                    void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream) {
  if (!w_a_q || !w_b_q) return fail(KEMR_ERR_ARG, "scan_topk_gated: null weight array");
  return scan_topk_impl(q, Q, gal_a, gal_b, M, D, 0.0, 0.0, w_a_q, w_b_q, alpha, hit_rowptr, hit_col, hit_bonus,
                        max_hits_per_query, k, k_sel, eps, idx_base, out_score64, out_score32, out_idx, out_flags,
                        workspace, workspace_bytes, path, stream);
}

static int score_pairs_impl(const uint16_t* q, const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                            double w_a, double w_b, const double* wq_a, const double* wq_b, double alpha,
                            const int32_t* pair_q, const int64_t* pair_row, const double* pair_bonus, int64_t n_pairs,
                            double* out_score64, kemr_stream_t stream) {
  if (!q || !gal_a || !pair_q || !pair_row || !out_score64) return fail(KEMR_ERR_ARG, "score_pairs: null pointer");
  if (D <= 0 || D > kMaxD || M <= 0) return fail(KEMR_ERR_ARG, "score_pairs: bad D or M");
  if (n_pairs <= 0) return KEMR_OK;
  const int64_t blocks = std::min<int64_t>((n_pairs + 7) / 8, 148 * 8);
  score_pairs_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(q, gal_a, gal_b, M, D, w_a, w_b, wq_a, wq_b, alpha, pair_q,
                                                              pair_row, pair_bonus, n_pairs, out_score64);
  LAUNCH_CHECK("score_pairs_kernel");
  return KEMR_OK;
}

extern "C" int kemr_score_pairs(const uint16_t* q, const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                                double w_a, double w_b, double alpha, const int32_t* pair_q,
                                const int64_t* pair_row, const double* pair_bonus, int64_t n_pairs,
                                double* out_score64, kemr_stream_t stream) {
  return score_pairs_impl(q, gal_a, gal_b, M, D, w_a, w_b, nullptr, nullptr, alpha, pair_q, pair_row, pair_bonus, n_pairs,
                          out_score64, stream);
}

extern "C" int kemr_score_pairs_gated(const uint16_t* q, const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                                      const double* w_a_q, const double* w_b_q, double alpha, const int32_t* pair_q,
                                      const int64_t* pair_row, const double* pair_bonus, int64_t n_pairs,
                                      double* out_score64, kemr_stream_t stream) {
  if (!w_a_q || !w_b_q || !gal_b) return fail(KEMR_ERR_ARG, "score_pairs_gated: needs both weight arrays and two galleries");
  return score_pairs_impl(q, gal_a, gal_b, M, D, 0.0, 0.0, w_a_q, w_b_q, alpha, pair_q, pair_row, pair_bonus, n_pairs,
                          out_score64, stream);
}

// ----------------------------------------------------------------------------- rank counting
static int rank_count_impl(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                               int64_t M, int D, double w_a, double w_b, const double* wq_a, const double* wq_b, double alpha,
                               const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                               const double* t_score64, const int64_t* t_gidx, double eps, int64_t idx_base,
                               int64_t* out_count, int32_t* out_flags, void* workspace,
                               size_t workspace_bytes, int path, kemr_stream_t stream) {
  int rc = check_common(q, Q, gal_a, M, D);
  if (rc) return rc;
  if (!(alpha > 0.0)) return fail(KEMR_ERR_ARG, "alpha must be > 0 (alpha=%g)", alpha);
  if (!t_score64 || !t_gidx || !out_count || !out_flags) return fail(KEMR_ERR_ARG, "rank_count: null pointer");
  if (hit_rowptr && (!hit_col || !hit_bonus)) return fail(KEMR_ERR_ARG, "hit CSR needs col and bonus arrays");
  DevInfo dv;
  if ((rc = dev_info(&dv))) return rc;
  const int G = gal_b ? 2 : 1;
  ScanPlan pl;
  if ((wq_a != nullptr) != (wq_b != nullptr) || (wq_a && !gal_b)) return fail(KEMR_ERR_ARG, "per-query weights need both arrays and two galleries");
  if ((rc = make_plan(Q, M, D, G, 16, kModeCount, path, !wq_a && (float)w_a == (float)w_b, dv, &pl))) return rc;
  const int Qrows = pl.path == KEMR_PATH_MMA ? pl.mma.q_pad : Q;

  // carve: band_lo | band_hi | part_count | amb_counter | amb_q | amb_row
  unsigned char* p = reinterpret_cast<unsigned char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const size_t o_lo = take((size_t)Qrows * 4), o_hi = take((size_t)Qrows * 4);
  const size_t o_pc = take((size_t)pl.P * Qrows * 4), o_ctr = take(256);
  const size_t o_w0 = take(wq_a ? (size_t)Q * 4 : 0), o_w1 = take(wq_a ? (size_t)Q * 4 : 0);
  if (!workspace || workspace_bytes < off + 2 * 4096) return fail(KEMR_ERR_WORKSPACE, "rank_count needs at least %zu workspace bytes, got %zu", off + 2 * 4096, workspace_bytes);
  const size_t left = workspace_bytes - off;
  const unsigned int amb_cap = (unsigned int)std::min<size_t>((left - 512) / 8, 0x7fffffffu);
  const size_t o_aq = take((size_t)amb_cap * 4), o_ar = off;
  float* band_lo = reinterpret_cast<float*>(p + o_lo);
  float* band_hi = reinterpret_cast<float*>(p + o_hi);
  int32_t* part_count = reinterpret_cast<int32_t*>(p + o_pc);
  unsigned int* amb_counter = reinterpret_cast<unsigned int*>(p + o_ctr);
  uint32_t* amb_q = reinterpret_cast<uint32_t*>(p + o_aq);
  uint32_t* amb_row = reinterpret_cast<uint32_t*>(p + o_ar);
  cudaStream_t st = S(stream);

  float* wq32[2] = {nullptr, nullptr};
  if (wq_a) {
    wq32[0] = reinterpret_cast<float*>(p + o_w0); wq32[1] = reinterpret_cast<float*>(p + o_w1);
    weights_to_f32_kernel<<<(Q + 255) / 256, 256, 0, st>>>(wq_a, wq_b, Q, wq32[0], wq32[1]);
    LAUNCH_CHECK("weights_to_f32_kernel");
  }
  CUDA_TRY(cudaMemsetAsync(amb_counter, 0, 256, st));
  if (Qrows > Q) {   // padded query rows of the tensor path never count
    CUDA_TRY(cudaMemsetAsync(band_lo, 0x7f, (size_t)Qrows * 4, st));   // ~3.4e38
    CUDA_TRY(cudaMemsetAsync(band_hi, 0x7f, (size_t)Qrows * 4, st));
  }
  rank_band_kernel<<<(Q + 255) / 256, 256, 0, st>>>(t_score64, alpha, eps, Q, band_lo, band_hi);
  LAUNCH_CHECK("rank_band_kernel");

  ScanArgs a{};
  a.q = q; a.Q = Q; a.gal[0] = gal_a; a.gal[1] = gal_b; a.G = G; a.M = M; a.D = D;
  a.w[0] = (float)w_a; a.w[1] = (float)w_b; a.wq[0] = wq32[0]; a.wq[1] = wq32[1]; a.mode = kModeCount; a.K = 16;
  a.band_lo = band_lo; a.band_hi = band_hi; a.part_count = part_count;
  a.amb_q = amb_q; a.amb_row = amb_row; a.amb_counter = amb_counter; a.amb_cap = amb_cap;
  if (pl.path == KEMR_PATH_MMA) {
    CUDA_TRY(cudaMemsetAsync(part_count, 0, (size_t)pl.P * Qrows * 4, st));
    if ((rc = mma_launch(a, pl.mma, st))) return fail(KEMR_ERR_CUDA, "tcgen05 scan launch failed: %s", mma_last_error());
  } else {
    if ((rc = launch_warp_scan(a, pl, st))) return rc;
  }
  if (g_scan_done_event) CUDA_TRY(cudaEventRecord(g_scan_done_event, st));

  unsigned long long* count = reinterpret_cast<unsigned long long*>(out_count);
  rank_sum_parts_kernel<<<(Q + 255) / 256, 256, 0, st>>>(part_count, pl.P, Q, Qrows, amb_counter, amb_cap, count, out_flags);
  LAUNCH_CHECK("rank_sum_parts_kernel");
  RankFixArgs f{};
  f.q = q; f.gal[0] = gal_a; f.gal[1] = gal_b; f.G = G; f.D = D; f.w[0] = w_a; f.w[1] = w_b; f.wq[0] = wq_a; f.wq[1] = wq_b; f.alpha = alpha;
  f.t = t_score64; f.t_gidx = t_gidx; f.idx_base = idx_base; f.count = count;
  rank_amb_kernel<<<dv.sms * 4, 256, 0, st>>>(f, amb_q, amb_row, amb_counter, amb_cap);
  LAUNCH_CHECK("rank_amb_kernel");
  if (hit_rowptr) {
    rank_hits_kernel<<<std::min(Q, dv.sms * 8), 128, 0, st>>>(f, Q, M, hit_rowptr, hit_col, hit_bonus);
    LAUNCH_CHECK("rank_hits_kernel");
  }
  return KEMR_OK;
}

extern "C" int kemr_rank_count(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                               int64_t M, int D, double w_a, double w_b, double alpha,
                               const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                               const double* t_score64, const int64_t* t_gidx, double eps, int64_t idx_base,
                               int64_t* out_count, int32_t* out_flags, void* workspace,
                               size_t workspace_bytes, int path, kemr_stream_t stream) {
  return rank_count_impl(q, Q, gal_a, gal_b, M, D, w_a, w_b, nullptr, nullptr, alpha, hit_rowptr, hit_col, hit_bonus,
                         t_score64, t_gidx, eps, idx_base, out_count, out_flags, workspace, workspace_bytes, path, stream);
}

extern "C" int kemr_rank_count_gated(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                                     int64_t M, int D, const double* w_a_q, const double* w_b_q, double alpha,
                                     const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                                     const double* t_score64, const int64_t* t_gidx, double eps, int64_t idx_base,
                                     int64_t* out_count, int32_t* out_flags, void* workspace,
                                     size_t workspace_bytes, int path, kemr_stream_t stream) {
  if (!w_a_q || !w_b_q) return fail(KEMR_ERR_ARG, "rank_count_gated: null weight array");
  return rank_count_impl(q, Q, gal_a, gal_b, M, D, 0.0, 0.0, w_a_q, w_b_q, alpha, hit_rowptr, hit_col, hit_bonus,
                         t_score64, t_gidx, eps, idx_base, out_count, out_flags, workspace, workspace_bytes, path, stream);
}

// ----------------------------------------------------------------------------- gate of the simple gated heads
// gate[q] = sigmoid(sum_d q[q][d] * weight[d] + bias) in fp32 (fusion_model.py:18-19, :190-191); one warp per query,
// lane-strided partial sums then a 16-8-4-2-1 fold.  w_a = gate, w_b = fl32(1 - gate), both widened to binary64.
__global__ void gate_linear_kernel(const uint16_t* __restrict__ q, int Q, int D, const float* __restrict__ weight,
                                   float bias, double* __restrict__ wa, double* __restrict__ wb) {
  const int lane = threadIdx.x & 31;
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= Q) return;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc = fmaf(bf16_to_f32(q[(size_t)qi * D + d]), weight[d], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float logit = acc + bias;
    const float gate = 1.0f / (1.0f + expf(-logit));
    wa[qi] = (double)gate;
    wb[qi] = (double)(1.0f - gate);
  }
}

extern "C" int kemr_gate_linear(const uint16_t* q, int Q, int D, const float* weight, float bias,
                                double* out_w_a_q, double* out_w_b_q, kemr_stream_t stream) {
  if (!q || !weight || !out_w_a_q || !out_w_b_q || Q <= 0 || D <= 0) return fail(KEMR_ERR_ARG, "gate_linear: bad argument");
  gate_linear_kernel<<<(Q + 7) / 8, 256, 0, S(stream)>>>(q, Q, D, weight, bias, out_w_a_q, out_w_b_q);
  LAUNCH_CHECK("gate_linear_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- dense score matrix
extern "C" int kemr_score_matrix(const uint16_t* q, int Q, const uint16_t* gal_a, const uint16_t* gal_b,
                                 int64_t M, int D, float w_a, float w_b, float* out, int64_t ld,
                                 void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  int rc = check_common(q, Q, gal_a, M, D);
  if (rc) return rc;
  if (!out || ld < M) return fail(KEMR_ERR_ARG, "score_matrix: bad output");
  DevInfo dv;
  if ((rc = dev_info(&dv))) return rc;
  const int G = gal_b ? 2 : 1;
  ScanPlan pl;
  if ((rc = make_plan(Q, M, D, G, 16, kModeDense, path, w_a == w_b, dv, &pl))) return rc;
  ScanArgs a{};
  a.q = q; a.Q = Q; a.gal[0] = gal_a; a.gal[1] = gal_b; a.G = G; a.M = M; a.D = D;
  a.w[0] = w_a; a.w[1] = w_b; a.wq[0] = nullptr; a.wq[1] = nullptr; a.mode = kModeDense; a.K = 16; a.dense = out; a.ld = ld;
  if (pl.path == KEMR_PATH_MMA) {
    if ((rc = mma_launch(a, pl.mma, S(stream)))) return fail(KEMR_ERR_CUDA, "tcgen05 scan launch failed: %s", mma_last_error());
    return KEMR_OK;
  }
  return launch_warp_scan(a, pl, S(stream));
}

// ----------------------------------------------------------------------------- matrix compat
extern "C" int kemr_matrix_rank(const float* Smat, int Q, int64_t M, int64_t ld, const int64_t* target_col,
                                int64_t* out_rank, kemr_stream_t stream) {
  if (!Smat || !target_col || !out_rank || Q <= 0 || M <= 0 || ld < M) return fail(KEMR_ERR_ARG, "matrix_rank: bad argument");
  matrix_rank_kernel<<<Q, 256, 0, S(stream)>>>(Smat, Q, M, ld, target_col, out_rank);
  LAUNCH_CHECK("matrix_rank_kernel");
  return KEMR_OK;
}

extern "C" int kemr_matrix_topk(const float* Smat, int Q, int64_t M, int64_t ld, int k, int64_t* out_idx,
                                float* out_val, kemr_stream_t stream) {
  if (!Smat || !out_idx || !out_val || Q <= 0 || M <= 0 || ld < M) return fail(KEMR_ERR_ARG, "matrix_topk: bad argument");
  if (k <= 0 || k > kMaxKSel) return fail(KEMR_ERR_ARG, "matrix_topk: k must be in [1, %d]", kMaxKSel);
  if (M >= (1ll << 32) - 1) return fail(KEMR_ERR_ARG, "matrix_topk: too many columns");
  matrix_topk_kernel<<<Q, 256, (size_t)8 * k * 8, S(stream)>>>(Smat, Q, M, ld, k, out_idx, out_val);
  LAUNCH_CHECK("matrix_topk_kernel");
  return KEMR_OK;
}

extern "C" int kemr_matrix_fuse(const float* Smat, float* out, int Q, int64_t M, int64_t ld, int scale_first,
                                float alpha32, const int64_t* hit_rowptr, const int32_t* hit_col,
                                const float* hit_add, kemr_stream_t stream) {
  if (!Smat || !out || Smat == out || Q <= 0 || M <= 0 || ld < M) return fail(KEMR_ERR_ARG, "matrix_fuse: bad argument");
  const int64_t total = (int64_t)Q * M;
  const int64_t blocks = std::min<int64_t>((total + 255) / 256, 148 * 16);
  matrix_scale_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(Smat, out, Q, M, ld, scale_first, alpha32);
  LAUNCH_CHECK("matrix_scale_kernel");
  if (hit_rowptr) {
    if (!hit_col || !hit_add) return fail(KEMR_ERR_ARG, "matrix_fuse: hit CSR needs col and add arrays");
    matrix_hits_kernel<<<(Q + 127) / 128, 128, 0, S(stream)>>>(out, Q, M, ld, hit_rowptr, hit_col, hit_add);
    LAUNCH_CHECK("matrix_hits_kernel");
  }
  return KEMR_OK;
}

extern "C" int kemr_matrix_mlp2(const float* S_a, const float* S_b, float* out, int Q, int64_t M, const float* w1,
                                const float* b1, const float* w2, float b2, int hidden, kemr_stream_t stream) {
  if (!S_a || !S_b || !out || !w1 || !b1 || !w2 || Q <= 0 || M <= 0 || hidden <= 0 || hidden > 2048)
    return fail(KEMR_ERR_ARG, "matrix_mlp2: bad argument");
  const int64_t total = (int64_t)Q * M;
  const int64_t blocks = std::min<int64_t>((total + 255) / 256, 148 * 8);
  matrix_mlp2_kernel<<<(unsigned)blocks, 256, (size_t)hidden * 16, S(stream)>>>(S_a, S_b, out, total, w1, b1, w2, b2, hidden);
  LAUNCH_CHECK("matrix_mlp2_kernel");
  return KEMR_OK;
}

extern "C" int kemr_infonce_rows(const float* a, const float* b, int B, int D, float temperature, float* out_row_loss,
                                 kemr_stream_t stream) {
  if (!a || !b || !out_row_loss || B <= 0 || D <= 0 || D % 4 || D > 1024 || !(temperature > 0.f))
    return fail(KEMR_ERR_ARG, "infonce_rows: bad argument (D %% 4 == 0, D <= 1024, temperature > 0)");
  if (((uintptr_t)a | (uintptr_t)b) & 15) return fail(KEMR_ERR_ARG, "infonce_rows: feature pointers must be 16-byte aligned");
  const int blocks = (B + 7) / 8;
  const float inv_tau = 1.0f / temperature;
  const int ch = (D + 127) / 128;
  cudaStream_t st = S(stream);
  switch (ch) {
#define KEMR_NCE(n) case n: infonce_rows_kernel<n><<<blocks, 256, 0, st>>>(a, b, B, D, inv_tau, out_row_loss); break;
    KEMR_NCE(1) KEMR_NCE(2) KEMR_NCE(3) KEMR_NCE(4) KEMR_NCE(5) KEMR_NCE(6) KEMR_NCE(7) KEMR_NCE(8)
#undef KEMR_NCE
  }
  LAUNCH_CHECK("infonce_rows_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- metrics reduction
struct MetricsScratch { double* sums = nullptr; int64_t* off = nullptr; int64_t* n = nullptr; int dev = -1; };
static int metrics_scratch(MetricsScratch** out) {
  static thread_local MetricsScratch sc;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!sc.sums || sc.dev != dev) {
    CUDA_TRY(cudaMalloc(&sc.sums, (size_t)kMaxLeaves * 8));
    CUDA_TRY(cudaMalloc(&sc.off, (size_t)kMaxLeaves * 8));
    CUDA_TRY(cudaMalloc(&sc.n, (size_t)kMaxLeaves * 8));
    sc.dev = dev;
  }
  *out = &sc;
  return KEMR_OK;
}

extern "C" int kemr_metrics_reduce(const int64_t* ranks, int Q, const int32_t* k_values, int n_k,
                                   int64_t* out_hits, double* out_stats, kemr_stream_t stream) {
  if (!ranks || Q <= 0 || n_k < 0 || n_k > 32 || !out_stats || (n_k && (!k_values || !out_hits)))
    return fail(KEMR_ERR_ARG, "metrics_reduce: bad argument");
  if ((int64_t)Q > (int64_t)kMaxLeaves * 64) return fail(KEMR_ERR_ARG, "metrics_reduce: Q too large");
  MetricsScratch* sc;
  int rc = metrics_scratch(&sc);
  if (rc) return rc;
  metrics_reduce_kernel<<<1, kMetricsThreads, 0, S(stream)>>>(ranks, Q, k_values, n_k, out_hits, out_stats,
                                                              sc->sums, sc->off, sc->n);
  LAUNCH_CHECK("metrics_reduce_kernel");
  return KEMR_OK;
}

extern "C" int kemr_metrics_reduce_host(const int64_t* ranks, int Q, const int32_t* k_values, int n_k,
                                        int64_t* out_hits, double* out_stats) {
  if (!ranks || Q <= 0 || n_k < 0 || n_k > 32 || !out_stats || (n_k && (!k_values || !out_hits)))
    return fail(KEMR_ERR_ARG, "metrics_reduce_host: bad argument");
  unsigned long long sum = 0;
  for (int j = 0; j < n_k; ++j) out_hits[j] = 0;
  for (int i = 0; i < Q; ++i) {
    sum += (unsigned long long)rank_pos(ranks[i]);
    for (int j = 0; j < n_k; ++j) out_hits[j] += ranks[i] > 0 && ranks[i] <= k_values[j];
  }
  out_stats[0] = (double)sum;
  out_stats[1] = 0.0 + pairwise_tree(Q, [&](int64_t off, int64_t n) { return pairwise_leaf(ranks, off, n); });
  return KEMR_OK;
}

extern "C" int kemr_merge_topk_strided(const double* in_score64, const int64_t* in_idx, int64_t rank_stride, int R, int Q, int k,
                                       double* out_score64, int64_t* out_idx, kemr_stream_t stream);
extern "C" int kemr_merge_topk(const double* in_score64, const int64_t* in_idx, int R, int Q, int k,
                               double* out_score64, int64_t* out_idx, kemr_stream_t stream) {
  return kemr_merge_topk_strided(in_score64, in_idx, (int64_t)Q * k, R, Q, k, out_score64, out_idx, stream);
}

extern "C" int kemr_merge_topk_strided(const double* in_score64, const int64_t* in_idx, int64_t rank_stride, int R, int Q, int k,
                                       double* out_score64, int64_t* out_idx, kemr_stream_t stream) {
  if (!in_score64 || !in_idx || !out_score64 || !out_idx || R <= 0 || Q <= 0 || k <= 0 || rank_stride < (int64_t)Q * k)
    return fail(KEMR_ERR_ARG, "merge_topk: bad argument");
  const size_t smem = (size_t)R * k * 16;
  if (smem > 200 * 1024 || R > 64) return fail(KEMR_ERR_ARG, "merge_topk: R*k too large (%d*%d)", R, k);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<Q, 256, smem, S(stream)>>>(in_score64, in_idx, (long long)rank_stride, R, Q, k, out_score64, out_idx, nullptr, 0, nullptr);
  LAUNCH_CHECK("merge_topk_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- result exchange over NVLink peer memory
// One exchange buffer per rank (cudaMalloc, exported with cudaIpcGetMemHandle or shared by pointer inside one process):
//   scores  f64 [2 parity][world][max_q * max_k]
//   indices i64 [2 parity][world][max_q * max_k]
//   flags   u32 [2 parity][world][max_q]          epoch of the step whose rows are complete
//   epoch   u32                                   bumped once per step by kemr_peer_begin
// A step of rank r: kemr_peer_begin -> kemr_scan_topk (its selection kernel stores the k rows of every query into
// slot r of EVERY rank's buffer and releases the query's flag there) -> kemr_peer_merge (one CTA per query waits for
// the world's flags of this epoch in LOCAL memory and merges).  Buffers alternate by epoch parity: a rank can only
// be one step ahead of the slowest one, because its next merge needs that rank's next rows.
struct kemr_peer {
  int rank = 0, world = 1, max_q = 0, max_k = 0, dev = -1;
  unsigned char* local = nullptr;
  unsigned char* base[kMaxPeers] = {nullptr};
  bool opened[kMaxPeers] = {false};
  size_t bytes = 0, idx_region = 0, flag_region = 0, epoch_off = 0;
};
static thread_local kemr_peer* g_push_peer = nullptr;      // armed by kemr_peer_begin, consumed by the next kemr_scan_topk

__global__ void peer_bump_kernel(unsigned int* epoch) { *epoch += 1u; }

// the selection kernel of the kemr_scan_topk call that follows kemr_peer_begin also stores its rows into every rank's
// buffer (select.cuh, stage D)
static int arm_peer_push(SelectArgs* s, int Q, int k) {
  kemr_peer* p = g_push_peer;
  s->n_peer = 0;
  if (!p) return KEMR_OK;
  g_push_peer = nullptr;
  if (Q > p->max_q || k > p->max_k) return fail(KEMR_ERR_ARG, "peer exchange: Q=%d / k=%d beyond the handle's limits (%d / %d)", Q, k, p->max_q, p->max_k);
  s->n_peer = p->world;
  for (int r = 0; r < p->world; ++r) s->peer_base[r] = p->base[r];
  s->peer_block = (long long)p->max_q * p->max_k;
  s->peer_idx_region = (long long)p->idx_region;
  s->peer_flag_region = (long long)p->flag_region;
  s->peer_max_q = p->max_q; s->peer_rank = p->rank; s->peer_world = p->world;
  s->peer_epoch = reinterpret_cast<const unsigned int*>(p->local + p->epoch_off);
  return KEMR_OK;
}

extern "C" int kemr_peer_create(int rank, int world, int max_queries, int max_k, kemr_peer_t** out, void* ipc_handle_host64) {
  if (!out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || max_queries <= 0 || max_k <= 0)
    return fail(KEMR_ERR_ARG, "peer_create: bad argument (world <= %d)", kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  kemr_peer* p = new kemr_peer();
  p->rank = rank; p->world = world; p->max_q = max_queries; p->max_k = max_k;
  const size_t block = (size_t)max_queries * max_k;
  p->idx_region = align_up((size_t)2 * world * block * 8);
  p->flag_region = p->idx_region + align_up((size_t)2 * world * block * 8);
  p->epoch_off = p->flag_region + align_up((size_t)2 * world * max_queries * 4);
  p->bytes = p->epoch_off + 256;
  cudaError_t e = cudaGetDevice(&p->dev);
  if (e == cudaSuccess) e = cudaMalloc(&p->local, p->bytes);
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->bytes);
  if (e == cudaSuccess && ipc_handle_host64) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_host64), p->local);
  if (e != cudaSuccess) { cudaFree(p->local); delete p; return fail(KEMR_ERR_CUDA, "peer_create: %s", cudaGetErrorString(e)); }
  p->base[rank] = p->local;
  *out = p;
  return KEMR_OK;
}

extern "C" int kemr_peer_connect(kemr_peer_t* p, const void* ipc_handles_host) {
  if (!p || !ipc_handles_host) return fail(KEMR_ERR_ARG, "peer_connect: null argument");
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(ipc_handles_host);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(KEMR_ERR_CUDA, "peer_connect: cannot map the buffer of rank %d: %s", r, cudaGetErrorString(e));
    p->base[r] = reinterpret_cast<unsigned char*>(ptr);
    p->opened[r] = true;
  }
  return KEMR_OK;
}

extern "C" int kemr_peer_connect_pointers(kemr_peer_t* p, void* const* bases_host) {
  if (!p || !bases_host) return fail(KEMR_ERR_ARG, "peer_connect_pointers: null argument");
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    if (!bases_host[r]) return fail(KEMR_ERR_ARG, "peer_connect_pointers: no buffer for rank %d", r);
    p->base[r] = reinterpret_cast<unsigned char*>(bases_host[r]);
  }
  return KEMR_OK;
}

extern "C" void* kemr_peer_local_buffer(kemr_peer_t* p) { return p ? p->local : nullptr; }

extern "C" int kemr_peer_destroy(kemr_peer_t* p) {
  if (!p) return KEMR_OK;
  if (g_push_peer == p) g_push_peer = nullptr;
  cudaDeviceSynchronize();
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->base[r]);
  cudaFree(p->local);
  delete p;
  return KEMR_OK;
}

extern "C" int kemr_peer_begin(kemr_peer_t* p, kemr_stream_t stream) {
  if (!p) return fail(KEMR_ERR_ARG, "peer_begin: null handle");
  for (int r = 0; r < p->world; ++r)
    if (!p->base[r]) return fail(KEMR_ERR_ARG, "peer_begin: rank %d is not connected", r);
  peer_bump_kernel<<<1, 1, 0, S(stream)>>>(reinterpret_cast<unsigned int*>(p->local + p->epoch_off));
  LAUNCH_CHECK("peer_bump_kernel");
  g_push_peer = p;
  return KEMR_OK;
}

extern "C" int kemr_peer_gather(kemr_peer_t* p, int Q, int k, double* out_score64, int64_t* out_idx, kemr_stream_t stream) {
  if (!p || !out_score64 || !out_idx || Q <= 0 || k <= 0 || Q > p->max_q || k > p->max_k)
    return fail(KEMR_ERR_ARG, "peer_gather: bad argument (Q <= %d, k <= %d)", p ? p->max_q : 0, p ? p->max_k : 0);
  if (g_push_peer == p) return fail(KEMR_ERR_ARG, "peer_gather: no kemr_scan_topk call since kemr_peer_begin");
  const size_t block = (size_t)p->max_q * p->max_k;
  gather_peer_kernel<<<dim3(Q, p->world), 32, 0, S(stream)>>>(p->local, (long long)p->idx_region, (long long)p->flag_region,
                                                              (long long)block, p->world, p->max_q, Q, k, out_score64, out_idx,
                                                              reinterpret_cast<const unsigned int*>(p->local + p->epoch_off));
  LAUNCH_CHECK("gather_peer_kernel");
  return KEMR_OK;
}

extern "C" int kemr_peer_merge(kemr_peer_t* p, int Q, int k, double* out_score64, int64_t* out_idx, kemr_stream_t stream) {
  if (!p || !out_score64 || !out_idx || Q <= 0 || k <= 0 || Q > p->max_q || k > p->max_k)
    return fail(KEMR_ERR_ARG, "peer_merge: bad argument (Q <= %d, k <= %d)", p ? p->max_q : 0, p ? p->max_k : 0);
  if (g_push_peer == p) return fail(KEMR_ERR_ARG, "peer_merge: no kemr_scan_topk call since kemr_peer_begin");
  const size_t smem = (size_t)p->world * k * 16;
  if (smem > 200 * 1024) return fail(KEMR_ERR_ARG, "peer_merge: world*k too large");
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(merge_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // both parities are passed as one base: the kernel picks the epoch's half through the rank stride arithmetic below
  const size_t block = (size_t)p->max_q * p->max_k;
  merge_peer_kernel<<<Q, 256, smem, S(stream)>>>(p->local, (long long)p->idx_region, (long long)p->flag_region, (long long)block,
                                                 p->world, p->max_q, Q, k, out_score64, out_idx,
                                                 reinterpret_cast<const unsigned int*>(p->local + p->epoch_off));
  LAUNCH_CHECK("merge_peer_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- resident index, host-buffer search
constexpr int kSmallBatch = 2;            // host-buffer searches up to this many queries take the one-launch route (the fused streaming
                                          // kernel wins up to two queries; from three on the tcgen05 scan does: profiles/r02_sweep_batch.jsonl)
constexpr int kMaxChunks = 4;             // large batches: query chunks whose PCIe transfer overlaps the previous chunk's scan
constexpr int kChunkUnit = 256;           // chunk boundaries sit on the tcgen05 scan's query blocks (one CTA pair = 256 queries)

// Chunk sizes of a large host-buffer batch whose queries sit in PAGEABLE memory: the staging memcpy of chunk i+1 runs
// on the calling thread while the copy engine moves chunk i and the SMs scan chunk i-1 (1000 queries: 411 -> 388 us).
// For page-locked caller buffers chunking was measured and rejected: only the first chunk's transfer stays exposed
// (57 -> 26 us at 1000 queries), but the scans of 256 + 744 queries take 115 us instead of 98 (fewer query blocks share
// a gallery stripe) and a second selection adds 10 us -- 216 us against 205 us for one quantise kernel reading the
// page-locked queries in place; three chunks: 250 us (profiles/r02_session_k_stdout.txt).  KEMR_E2E_CHUNKS="256,256"
// forces leading chunk sizes for either kind of buffer (experiments; "0" = never).
static int plan_chunks(int Q, bool page_locked, int* sizes) {
  static const char* env = getenv("KEMR_E2E_CHUNKS");
  int n = 0, left = Q;
  if (env) {
    const char* p = env;
    while (*p && n < kMaxChunks - 1) {
      const int v = atoi(p);
      if (v <= 0 || v >= left) break;
      sizes[n++] = v; left -= v;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
  } else if (!page_locked && Q >= 2 * kChunkUnit) {
    sizes[n++] = kChunkUnit; left -= kChunkUnit;
  }
  sizes[n++] = left;
  return n;
}

struct kemr_index {
  bool owns_gal = true;                         // false: a lane created by kemr_index_share (the galleries belong to the source)
  // a submitted search that has not been waited for: what kemr_index_wait still has to do
  int pending = 0;                              // 0 none, 1 synchronise only (results already land in the caller's arrays),
                                                // 2 + copy the handle's page-locked result arrays out, 3 + device-to-host copies first
  int p_Q = 0, p_k = 0;
  int64_t* p_idx = nullptr; double* p_score = nullptr; int32_t* p_flags = nullptr;
  uint16_t* gal[2] = {nullptr, nullptr};
  float gal_norm[2] = {1.f, 1.f};               // largest row norm of each resident gallery (measured once at creation)
  int64_t M = 0;
  int D = 0, max_q = 0, max_k = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;           // query chunks travel here while the previous chunk is scanned
  cudaEvent_t chunk_ev[kMaxChunks] = {};
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // device staging
  float* d_qf32 = nullptr; uint16_t* d_q = nullptr;
  double* d_score = nullptr; int64_t* d_idx = nullptr; int32_t* d_flags = nullptr;
  int64_t* d_rowptr = nullptr; int32_t* d_col = nullptr; double* d_bonus = nullptr; int64_t hit_cap = 0;
  // one request blob for small batches: [q fp32 | hit rowptr | hit bonus | hit col] -> ONE host-to-device copy
  unsigned char* h_blob = nullptr; unsigned char* d_blob = nullptr; size_t blob_bytes = 0;
  // pinned staging
  float* h_q = nullptr; double* h_score = nullptr; int64_t* h_idx = nullptr; int32_t* h_flags = nullptr;
  int64_t* h_rowptr = nullptr; int32_t* h_col = nullptr; double* h_bonus = nullptr;
};

extern "C" int kemr_index_destroy(kemr_index_t* ix) {
  if (!ix) return KEMR_OK;
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  if (ix->owns_gal) { cudaFree(ix->gal[0]); cudaFree(ix->gal[1]); }
  cudaFree(ix->ws);
  cudaFree(ix->d_qf32); cudaFree(ix->d_q); cudaFree(ix->d_score); cudaFree(ix->d_idx); cudaFree(ix->d_flags);
  cudaFree(ix->d_rowptr); cudaFree(ix->d_col); cudaFree(ix->d_bonus);
  cudaFreeHost(ix->h_q); cudaFreeHost(ix->h_score); cudaFreeHost(ix->h_idx); cudaFreeHost(ix->h_flags);
  cudaFreeHost(ix->h_rowptr); cudaFreeHost(ix->h_col); cudaFreeHost(ix->h_bonus);
  cudaFreeHost(ix->h_blob); cudaFree(ix->d_blob);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
  for (int i = 0; i < kMaxChunks; ++i) if (ix->chunk_ev[i]) cudaEventDestroy(ix->chunk_ev[i]);
  delete ix;
  return KEMR_OK;
}

// everything a handle needs besides the galleries: streams, workspace, device and page-locked staging
static int index_alloc_buffers(kemr_index* ix) {
  const int64_t M = ix->M; const int D = ix->D; const int max_queries = ix->max_q, max_k = ix->max_k;
  ix->hit_cap = (int64_t)max_queries * 256;
  const int ksel = std::min(kMaxKSel, (max_k + 6 + 7) / 8 * 8);
  ix->ws_bytes = kemr_workspace_bytes(max_queries, M, D, ksel, 256);
#define IX_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    return fail(KEMR_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); } while (0)
  IX_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  IX_TRY(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < kMaxChunks; ++i) IX_TRY(cudaEventCreateWithFlags(&ix->chunk_ev[i], cudaEventDisableTiming));
  IX_TRY(cudaMalloc(&ix->ws, ix->ws_bytes));
  const size_t nq = (size_t)max_queries;
  IX_TRY(cudaMalloc(&ix->d_qf32, nq * D * 4)); IX_TRY(cudaMalloc(&ix->d_q, nq * D * 2));
  IX_TRY(cudaMalloc(&ix->d_score, nq * max_k * 8)); IX_TRY(cudaMalloc(&ix->d_idx, nq * max_k * 8));
  IX_TRY(cudaMalloc(&ix->d_flags, nq * 4));
  IX_TRY(cudaMalloc(&ix->d_rowptr, (nq + 1) * 8)); IX_TRY(cudaMalloc(&ix->d_col, (size_t)ix->hit_cap * 4));
  IX_TRY(cudaMalloc(&ix->d_bonus, (size_t)ix->hit_cap * 8));
  IX_TRY(cudaMallocHost(&ix->h_q, nq * D * 4)); IX_TRY(cudaMallocHost(&ix->h_score, nq * max_k * 8));
  IX_TRY(cudaMallocHost(&ix->h_idx, nq * max_k * 8)); IX_TRY(cudaMallocHost(&ix->h_flags, nq * 4));
  IX_TRY(cudaMallocHost(&ix->h_rowptr, (nq + 1) * 8)); IX_TRY(cudaMallocHost(&ix->h_col, (size_t)ix->hit_cap * 4));
  IX_TRY(cudaMallocHost(&ix->h_bonus, (size_t)ix->hit_cap * 8));
  ix->blob_bytes = align_up((size_t)kSmallBatch * D * 4) + align_up((size_t)(kSmallBatch + 1) * 8) + (size_t)kSmallBatch * 256 * 12 + 256;
  IX_TRY(cudaMallocHost(&ix->h_blob, ix->blob_bytes)); IX_TRY(cudaMalloc(&ix->d_blob, ix->blob_bytes));
#undef IX_TRY
  return KEMR_OK;
}

extern "C" int kemr_index_create(const uint16_t* gal_a_host, const uint16_t* gal_b_host, int64_t M, int D,
                                 int max_queries, int max_k, kemr_index_t** out) {
  if (!gal_a_host || !out || M <= 0 || D <= 0 || D % 8 || D > kMaxD || max_queries <= 0 || max_k <= 0 || max_k > kMaxKSel - 8)
    return fail(KEMR_ERR_ARG, "index_create: bad argument");
  kemr_index* ix = new kemr_index();
  ix->M = M; ix->D = D; ix->max_q = max_queries; ix->max_k = max_k;
  const size_t gbytes = (size_t)M * D * 2;
  int rc = index_alloc_buffers(ix);
  if (rc) { kemr_index_destroy(ix); return rc; }
#define IX_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { kemr_index_destroy(ix); \
    return fail(KEMR_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); } } while (0)
  IX_TRY(cudaMalloc(&ix->gal[0], gbytes));
  IX_TRY(cudaMemcpyAsync(ix->gal[0], gal_a_host, gbytes, cudaMemcpyHostToDevice, ix->stream));
  if (gal_b_host) {
    IX_TRY(cudaMalloc(&ix->gal[1], gbytes));
    IX_TRY(cudaMemcpyAsync(ix->gal[1], gal_b_host, gbytes, cudaMemcpyHostToDevice, ix->stream));
  }
  // the selection margin of every search scales with the galleries' largest row norms (index_eps)
  float* d_norm = reinterpret_cast<float*>(ix->d_blob);
  for (int g = 0; g < 2; ++g) {
    if (!ix->gal[g]) continue;
    if ((rc = kemr_row_norm_max(ix->gal[g], M, D, d_norm + g, ix->stream))) { kemr_index_destroy(ix); return rc; }
    IX_TRY(cudaMemcpyAsync(&ix->gal_norm[g], d_norm + g, 4, cudaMemcpyDeviceToHost, ix->stream));
  }
  IX_TRY(cudaStreamSynchronize(ix->stream));
#undef IX_TRY
  if (!std::isfinite(ix->gal_norm[0]) || !std::isfinite(ix->gal_norm[1])) {
    kemr_index_destroy(ix);
    return fail(KEMR_ERR_ARG, "index_create: the galleries contain non-finite values");
  }
  *out = ix;
  return KEMR_OK;
}

// A second LANE over the same resident galleries: its own stream, workspace and staging buffers, so that a search
// submitted on one lane overlaps the transfers of the next batch on the other (kemr_index_submit_host / kemr_index_wait).
// The source handle must outlive its lanes.
extern "C" int kemr_index_share(kemr_index_t* src, kemr_index_t** out) {
  if (!src || !out) return fail(KEMR_ERR_ARG, "index_share: null argument");
  kemr_index* ix = new kemr_index();
  ix->owns_gal = false;
  ix->gal[0] = src->gal[0]; ix->gal[1] = src->gal[1];
  ix->M = src->M; ix->D = src->D; ix->max_q = src->max_q; ix->max_k = src->max_k;
  ix->gal_norm[0] = src->gal_norm[0]; ix->gal_norm[1] = src->gal_norm[1];
  int rc = index_alloc_buffers(ix);
  if (rc) { kemr_index_destroy(ix); return rc; }
  *out = ix;
  return KEMR_OK;
}

// Selection margin of a host-buffer search: the bound on |fp32 scan score - canonical score| is 2e-5 for ||q||, ||g|| <= 1
// and |w_a| + |w_b| <= 1 and scales with max||q|| * (|w_a| max||g_a|| + |w_b| max||g_b||) (engine.eps_for does the same
// for device-resident searches).  The galleries' norms are index state; the queries are unit rows after normalize = 1
// and must satisfy ||q|| <= 1 (up to bf16 rounding, as CLIP embeddings do) with normalize = 0 -- kemr.h says so.
static double index_eps(const kemr_index* ix, double w_a, double w_b) {
  double s = std::fabs(w_a) * ix->gal_norm[0];
  if (ix->gal[1]) s += std::fabs(w_b) * ix->gal_norm[1];
  return 2e-5 * std::max(1.0, s * (1.0 + 1.0 / 128.0));
}

// what a submitted search still owes its caller: wait for the stream, bring the results home
static int index_finish(kemr_index_t* ix) {
  if (!ix->pending) return KEMR_OK;
  const int kind = ix->pending;
  ix->pending = 0;
  cudaStream_t st = ix->stream;
  const size_t n = (size_t)ix->p_Q * ix->p_k;
  if (kind == 3) {
    CUDA_TRY(cudaMemcpyAsync(ix->h_idx, ix->d_idx, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ix->h_score, ix->d_score, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ix->h_flags, ix->d_flags, (size_t)ix->p_Q * 4, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (kind >= 2) {
    memcpy(ix->p_idx, ix->h_idx, n * 8);
    memcpy(ix->p_score, ix->h_score, n * 8);
    if (ix->p_flags) memcpy(ix->p_flags, ix->h_flags, (size_t)ix->p_Q * 4);
  }
  return KEMR_OK;
}

// q_bf16 != null: the queries arrive as bf16 bit patterns (half the bytes over PCIe, no quantise kernel)
static int index_search_host_impl(kemr_index_t* ix, const float* q_host, const uint16_t* q_bf16, int Q, int normalize,
                                  double w_a, double w_b, double alpha, const int64_t* hit_rowptr_host,
                                  const int32_t* hit_col_host, const double* hit_bonus_host, int k,
                                  int64_t* out_idx_host, double* out_score64_host, int32_t* out_flags_host, bool wait = true) {
  if (!ix || (!q_host && !q_bf16) || !out_idx_host || !out_score64_host) return fail(KEMR_ERR_ARG, "index_search_host: null pointer");
  if (ix->pending) return fail(KEMR_ERR_ARG, "index_search_host: a submitted search is still pending on this handle (kemr_index_wait)");
  if (Q <= 0 || Q > ix->max_q || k <= 0 || k > ix->max_k) return fail(KEMR_ERR_ARG, "index_search_host: Q or k beyond the handle's limits");
  cudaStream_t st = ix->stream;
  const double eps = index_eps(ix, w_a, w_b);
  const size_t qbytes = (size_t)Q * ix->D * 4;
  // Page-locked caller buffers are used IN PLACE by the kernels (the quantise kernel reads the fp32 queries over
  // PCIe, the select kernel stores the results straight into the caller's arrays): no copy engine hop in either
  // direction.  Pageable buffers are staged through the handle's page-locked memory.
  auto mapped = [](const void* p) -> void* {
    // page-locked host memory: its device view comes with the attributes (one driver call, not two)
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost) return nullptr;
    return at.devicePointer;
  };
  static const bool no_zero_copy = getenv("KEMR_NO_ZERO_COPY") != nullptr;
  if (Q <= kSmallBatch && !q_bf16) {
    // Serving-size batches: everything the request brings (queries, KG-hit CSR) goes over in ONE copy, and ONE kernel
    // does the rest (quantise, scan, selection); results come back with one copy per array.  2 + 3 copies and two
    // kernels cost 119 us for a batch of one; the scan itself is ~35 us.
    int64_t nnz = 0, max_hits = 0;
    if (hit_rowptr_host) {
      nnz = hit_rowptr_host[Q];
      if (nnz > (int64_t)kSmallBatch * 256) return fail(KEMR_ERR_ARG, "index_search_host: too many KG hits (%lld)", (long long)nnz);
      for (int i = 0; i < Q; ++i) max_hits = std::max(max_hits, hit_rowptr_host[i + 1] - hit_rowptr_host[i]);
    }
    const size_t o_rp = align_up(qbytes), o_bo = o_rp + align_up((size_t)(Q + 1) * 8), o_co = o_bo + align_up((size_t)nnz * 8);
    const size_t total = o_co + align_up((size_t)nnz * 4);
    // (Measured and rejected: the request as a kernel PARAMETER of 8 KB instead of this copy -- 63.1 vs 62.6 us per
    // call, the copy's DMA latency and the larger launch cancel; profiles/r02_session_k_stdout.txt.  Also rejected: a
    // completion word in page-locked memory that the kernel writes last and the host spins on instead of the stream
    // synchronise below -- 63.1 us with the spin, 61.1 us with the synchronise; profiles/r02_session_o_stdout.txt.)
    memcpy(ix->h_blob, q_host, qbytes);
    if (hit_rowptr_host) {
      memcpy(ix->h_blob + o_rp, hit_rowptr_host, (size_t)(Q + 1) * 8);
      memcpy(ix->h_blob + o_bo, hit_bonus_host, (size_t)nnz * 8);
      memcpy(ix->h_blob + o_co, hit_col_host, (size_t)nnz * 4);
    }
    CUDA_TRY(cudaMemcpyAsync(ix->d_blob, ix->h_blob, total, cudaMemcpyHostToDevice, st));
    const int ksel = std::min(kMaxKSel, (k + 6 + 7) / 8 * 8);
    // results: the selection stage stores straight into page-locked host memory (the caller's arrays when they are
    // page-locked, else the handle's staging arrays) -- no device-to-host copy, the stream synchronise is the fence
    int64_t* oi = static_cast<int64_t*>(mapped(out_idx_host));
    double* os = static_cast<double*>(mapped(out_score64_host));
    int32_t* of = out_flags_host ? static_cast<int32_t*>(mapped(out_flags_host)) : nullptr;
    const bool direct = !no_zero_copy && oi && os && (!out_flags_host || of);
    if (!direct) {
      oi = static_cast<int64_t*>(mapped(ix->h_idx)); os = static_cast<double*>(mapped(ix->h_score));
      of = static_cast<int32_t*>(mapped(ix->h_flags));
      if (!oi || !os || !of) return fail(KEMR_ERR_CUDA, "index_search_host: staging buffers are not mapped");
    } else if (!of) {
      of = static_cast<int32_t*>(mapped(ix->h_flags));
    }
    int rc = scan_topk_impl(ix->d_q, Q, ix->gal[0], ix->gal[1], ix->M, ix->D, w_a, w_b, nullptr, nullptr, alpha,
                            hit_rowptr_host ? reinterpret_cast<const int64_t*>(ix->d_blob + o_rp) : nullptr,
                            reinterpret_cast<const int32_t*>(ix->d_blob + o_co), reinterpret_cast<const double*>(ix->d_blob + o_bo),
                            max_hits, k, ksel, eps, 0, os, nullptr, oi, of, ix->ws, ix->ws_bytes,
                            KEMR_PATH_WARP, st, reinterpret_cast<const float*>(ix->d_blob), normalize);
    if (rc) return rc;
    ix->pending = direct ? 1 : 2;
    ix->p_Q = Q; ix->p_k = k; ix->p_idx = out_idx_host; ix->p_score = out_score64_host; ix->p_flags = out_flags_host;
    return wait ? index_finish(ix) : KEMR_OK;
  }
  const float* q_dev_view = (no_zero_copy || q_bf16) ? nullptr : static_cast<const float*>(mapped(q_host));
  int64_t* oi_view = no_zero_copy ? nullptr : static_cast<int64_t*>(mapped(out_idx_host));
  double* os_view = no_zero_copy ? nullptr : static_cast<double*>(mapped(out_score64_host));
  int32_t* of_view = (no_zero_copy || !out_flags_host) ? nullptr : static_cast<int32_t*>(mapped(out_flags_host));
  const bool q_pinned = q_dev_view != nullptr;
  const bool out_pinned = oi_view && os_view && (!out_flags_host || of_view);
  int64_t max_hits = 0;
  const int64_t* d_rowptr = nullptr;
  if (hit_rowptr_host) {
    const int64_t nnz = hit_rowptr_host[Q];
    if (nnz > ix->hit_cap) return fail(KEMR_ERR_ARG, "index_search_host: too many KG hits (%lld > %lld)", (long long)nnz, (long long)ix->hit_cap);
    for (int i = 0; i < Q; ++i) max_hits = std::max(max_hits, hit_rowptr_host[i + 1] - hit_rowptr_host[i]);
    memcpy(ix->h_rowptr, hit_rowptr_host, (size_t)(Q + 1) * 8);
    memcpy(ix->h_col, hit_col_host, (size_t)nnz * 4);
    memcpy(ix->h_bonus, hit_bonus_host, (size_t)nnz * 8);
    CUDA_TRY(cudaMemcpyAsync(ix->d_rowptr, ix->h_rowptr, (size_t)(Q + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnz) {
      CUDA_TRY(cudaMemcpyAsync(ix->d_col, ix->h_col, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(ix->d_bonus, ix->h_bonus, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    }
    d_rowptr = ix->d_rowptr;
  }
  const int ksel = std::min(kMaxKSel, (k + 6 + 7) / 8 * 8);
  double* os_dev = out_pinned ? os_view : ix->d_score;
  int64_t* oi_dev = out_pinned ? oi_view : ix->d_idx;
  int32_t* of_dev = (out_pinned && of_view) ? of_view : ix->d_flags;
  int rc = KEMR_OK;
  int sizes[kMaxChunks];
  const int nchunk = plan_chunks(Q, q_bf16 ? mapped(q_bf16) != nullptr : q_pinned, sizes);
  if (nchunk > 1) {
    // Pipelined: the copy engine brings chunk i+1 (page-locked caller memory is read in place, pageable memory goes
    // through the handle's page-locked buffer chunk by chunk) while the SMs quantise, scan and select chunk i; the
    // selection stores every chunk's rows straight into the caller's page-locked result arrays.
    const size_t esz = q_bf16 ? 2 : 4;                           // bytes per query element on the host
    const unsigned char* src = reinterpret_cast<const unsigned char*>(q_bf16 ? (const void*)q_bf16 : (const void*)q_host);
    const bool src_locked = mapped(src) != nullptr;
    unsigned char* dst = q_bf16 ? reinterpret_cast<unsigned char*>(ix->d_q) : reinterpret_cast<unsigned char*>(ix->d_qf32);
    int starts[kMaxChunks + 1];
    starts[0] = 0;
    for (int c = 0; c < nchunk; ++c) starts[c + 1] = starts[c] + sizes[c];
    auto send = [&](int c) -> cudaError_t {
      const size_t off = (size_t)starts[c] * ix->D * esz, bytes = (size_t)sizes[c] * ix->D * esz;
      const unsigned char* from = src + off;
      if (!src_locked) { from = reinterpret_cast<unsigned char*>(ix->h_q) + off; memcpy(const_cast<unsigned char*>(from), src + off, bytes); }
      cudaError_t e = cudaMemcpyAsync(dst + off, from, bytes, cudaMemcpyHostToDevice, ix->copy_stream);
      return e != cudaSuccess ? e : cudaEventRecord(ix->chunk_ev[c], ix->copy_stream);
    };
    // page-locked source: all transfers are queued at once (the copy engine runs them back to back); pageable source:
    // the staging memcpy of chunk i+1 runs on this thread while chunk i is in flight
    if (src_locked) for (int c = 0; c < nchunk; ++c) CUDA_TRY(send(c));
    for (int c = 0; c < nchunk; ++c) {
      const int q0 = starts[c], nq = sizes[c];
      if (!src_locked) CUDA_TRY(send(c));
      CUDA_TRY(cudaStreamWaitEvent(st, ix->chunk_ev[c], 0));
      uint16_t* qc = ix->d_q + (size_t)q0 * ix->D;
      if (!q_bf16 && (rc = kemr_quantize_rows(ix->d_qf32 + (size_t)q0 * ix->D, qc, nq, ix->D, normalize, st))) return rc;
      rc = kemr_scan_topk(qc, nq, ix->gal[0], ix->gal[1], ix->M, ix->D, w_a, w_b, alpha, d_rowptr ? d_rowptr + q0 : nullptr,
                          ix->d_col, ix->d_bonus, max_hits, k, ksel, eps, 0, os_dev + (size_t)q0 * k, nullptr,
                          oi_dev + (size_t)q0 * k, of_dev + q0, ix->ws, ix->ws_bytes, KEMR_PATH_AUTO, st);
      if (rc) return rc;
    }
  } else {
    if (q_bf16) {
      // bf16 bit patterns: one copy-engine transfer of Q*D*2 bytes straight into the scan's query buffer (page-locked
      // caller memory is read in place by the copy engine, pageable memory is staged)
      const void* src = q_bf16;
      if (!mapped(q_bf16)) { memcpy(ix->h_q, q_bf16, qbytes / 2); src = ix->h_q; }
      CUDA_TRY(cudaMemcpyAsync(ix->d_q, src, qbytes / 2, cudaMemcpyHostToDevice, st));
    } else if (!q_pinned) {
      memcpy(ix->h_q, q_host, qbytes);
      CUDA_TRY(cudaMemcpyAsync(ix->d_qf32, ix->h_q, qbytes, cudaMemcpyHostToDevice, st));
    } else if (!wait) {
      // submitted (pipelined) search: the COPY ENGINE brings the page-locked queries, so the transfer runs beside the
      // other lane's scan -- the quantise kernel reading them in place cannot: a scan CTA leaves an SM too few
      // registers for its blocks, and it queued behind the scan (C2: 180 us per step, the sum of transfer and compute)
      CUDA_TRY(cudaMemcpyAsync(ix->d_qf32, q_host, qbytes, cudaMemcpyHostToDevice, st));
    }
    const bool q_in_place = q_pinned && wait;
    rc = q_bf16 ? KEMR_OK : kemr_quantize_rows(q_in_place ? q_dev_view : ix->d_qf32, ix->d_q, Q, ix->D, normalize, st);
    if (rc) return rc;
    rc = kemr_scan_topk(ix->d_q, Q, ix->gal[0], ix->gal[1], ix->M, ix->D, w_a, w_b, alpha, d_rowptr, ix->d_col,
                        ix->d_bonus, max_hits, k, ksel, eps, 0, os_dev, nullptr, oi_dev, of_dev,
                        ix->ws, ix->ws_bytes, KEMR_PATH_AUTO, st);
  }
  if (rc) return rc;
  ix->pending = out_pinned ? 1 : 3;
  ix->p_Q = Q; ix->p_k = k; ix->p_idx = out_idx_host; ix->p_score = out_score64_host; ix->p_flags = out_flags_host;
  return wait ? index_finish(ix) : KEMR_OK;
}

// The same search without the wait: everything is queued on the handle's stream and the call returns; kemr_index_wait
// blocks until the results are in the caller's arrays.  One search per handle at a time -- a second lane
// (kemr_index_share) takes the next batch meanwhile, so its queries cross PCIe while this scan runs.  The caller's
// buffers (queries, hit CSR, results) must stay untouched until the wait returns when they are page-locked (they are
// used in place); pageable query / CSR buffers are copied before the call returns.
extern "C" int kemr_index_submit_host(kemr_index_t* ix, const float* q_host, int Q, int normalize,
                                      double w_a, double w_b, double alpha, const int64_t* hit_rowptr_host,
                                      const int32_t* hit_col_host, const double* hit_bonus_host, int k,
                                      int64_t* out_idx_host, double* out_score64_host, int32_t* out_flags_host) {
  if (!q_host) return fail(KEMR_ERR_ARG, "index_submit_host: null query pointer");
  return index_search_host_impl(ix, q_host, nullptr, Q, normalize, w_a, w_b, alpha, hit_rowptr_host, hit_col_host, hit_bonus_host, k,
                                out_idx_host, out_score64_host, out_flags_host, false);
}

extern "C" int kemr_index_wait(kemr_index_t* ix) {
  if (!ix) return fail(KEMR_ERR_ARG, "index_wait: null handle");
  return index_finish(ix);
}

extern "C" int kemr_index_search_host(kemr_index_t* ix, const float* q_host, int Q, int normalize,
                                      double w_a, double w_b, double alpha, const int64_t* hit_rowptr_host,
                                      const int32_t* hit_col_host, const double* hit_bonus_host, int k,
                                      int64_t* out_idx_host, double* out_score64_host, int32_t* out_flags_host) {
  if (!q_host) return fail(KEMR_ERR_ARG, "index_search_host: null query pointer");
  return index_search_host_impl(ix, q_host, nullptr, Q, normalize, w_a, w_b, alpha, hit_rowptr_host, hit_col_host, hit_bonus_host, k,
                                out_idx_host, out_score64_host, out_flags_host);
}

extern "C" int kemr_index_search_host_bf16(kemr_index_t* ix, const uint16_t* q_bf16_host, int Q,
                                           double w_a, double w_b, double alpha, const int64_t* hit_rowptr_host,
                                           const int32_t* hit_col_host, const double* hit_bonus_host, int k,
                                           int64_t* out_idx_host, double* out_score64_host, int32_t* out_flags_host) {
  if (!q_bf16_host) return fail(KEMR_ERR_ARG, "index_search_host_bf16: null query pointer");
  return index_search_host_impl(ix, nullptr, q_bf16_host, Q, 0, w_a, w_b, alpha, hit_rowptr_host, hit_col_host, hit_bonus_host, k,
                                out_idx_host, out_score64_host, out_flags_host);
}

// ----------------------------------------------------------------------------- KG-hit CSR builder (device)
extern "C" size_t kemr_hits_workspace_bytes(int Q) { return align_up((size_t)std::max(Q, 1) * 8) + 256; }

extern "C" int kemr_hits_build_csr(const int64_t* list_rowptr, const int64_t* list_rows, const double* bonus_per_query,
                                   int Q, int64_t row_lo, int64_t row_hi, int sum_repeats, int64_t* out_rowptr,
                                   int32_t* out_col, double* out_bonus, int64_t* out_max_per_query, void* workspace,
                                   size_t workspace_bytes, kemr_stream_t stream) {
  if (!list_rowptr || !bonus_per_query || !out_rowptr || !out_max_per_query || Q <= 0 || row_hi < row_lo)
    return fail(KEMR_ERR_ARG, "hits_build_csr: bad argument");
  if (row_hi - row_lo >= (1ll << 31)) return fail(KEMR_ERR_ARG, "hits_build_csr: a shard holds at most 2^31-1 rows");
  if (!workspace || workspace_bytes < kemr_hits_workspace_bytes(Q)) return fail(KEMR_ERR_WORKSPACE, "hits_build_csr: workspace too small");
  int64_t* count = reinterpret_cast<int64_t*>(workspace);
  cudaStream_t st = S(stream);
  const int blocks = (Q + kHitsWarpsPerBlock - 1) / kHitsWarpsPerBlock;
  hits_count_kernel<<<blocks, kHitsWarpsPerBlock * 32, 0, st>>>(list_rowptr, list_rows, Q, row_lo, row_hi, count);
  LAUNCH_CHECK("hits_count_kernel");
  hits_scan_kernel<<<1, 1024, 0, st>>>(count, Q, out_rowptr, out_max_per_query);
  LAUNCH_CHECK("hits_scan_kernel");
  hits_fill_kernel<<<blocks, kHitsWarpsPerBlock * 32, 0, st>>>(list_rowptr, list_rows, bonus_per_query, Q, row_lo, row_hi,
                                                                 sum_repeats, out_rowptr, out_col, out_bonus);
  LAUNCH_CHECK("hits_fill_kernel");
  return KEMR_OK;
}

extern "C" int kemr_hits_filter_csr(const int64_t* rowptr, const int32_t* col, const double* bonus, const int64_t* query_sel,
                                    int Q_out, int64_t col_lo, int64_t col_hi, int64_t* out_rowptr, int32_t* out_col,
                                    double* out_bonus, int64_t* out_max_per_query, void* workspace, size_t workspace_bytes,
                                    kemr_stream_t stream) {
  if (!rowptr || !out_rowptr || !out_max_per_query || Q_out <= 0 || col_hi < col_lo)
    return fail(KEMR_ERR_ARG, "hits_filter_csr: bad argument");
  if (!workspace || workspace_bytes < kemr_hits_workspace_bytes(Q_out)) return fail(KEMR_ERR_WORKSPACE, "hits_filter_csr: workspace too small");
  int64_t* count = reinterpret_cast<int64_t*>(workspace);
  cudaStream_t st = S(stream);
  const int blocks = (Q_out + kHitsWarpsPerBlock - 1) / kHitsWarpsPerBlock;
  hits_filter_count_kernel<<<blocks, kHitsWarpsPerBlock * 32, 0, st>>>(rowptr, col, query_sel, Q_out, col_lo, col_hi, count);
  LAUNCH_CHECK("hits_filter_count_kernel");
  hits_scan_kernel<<<1, 1024, 0, st>>>(count, Q_out, out_rowptr, out_max_per_query);
  LAUNCH_CHECK("hits_scan_kernel");
  hits_filter_fill_kernel<<<blocks, kHitsWarpsPerBlock * 32, 0, st>>>(rowptr, col, bonus, query_sel, Q_out, col_lo, col_hi,
                                                                        out_rowptr, out_col, out_bonus);
  LAUNCH_CHECK("hits_filter_fill_kernel");
  return KEMR_OK;
}

extern "C" int kemr_hits_target_bonus(const int64_t* rowptr, const int32_t* col, const double* bonus, int Q,
                                      const int64_t* target_col, double* out_bonus, kemr_stream_t stream) {
  if (!rowptr || !target_col || !out_bonus || Q <= 0) return fail(KEMR_ERR_ARG, "hits_target_bonus: bad argument");
  const int blocks = (Q + kHitsWarpsPerBlock - 1) / kHitsWarpsPerBlock;
  hits_target_bonus_kernel<<<blocks, kHitsWarpsPerBlock * 32, 0, S(stream)>>>(rowptr, col, bonus, Q, target_col, out_bonus);
  LAUNCH_CHECK("hits_target_bonus_kernel");
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- uuid -> row map (host)
struct kemr_idmap {
  std::string blob;                                        // owns the key bytes
  std::unordered_map<std::string_view, int64_t> map;
};

// last '/' segment of an artefact URI (fusion.py:76, text2sparql_retrieval.py:57)
static inline std::string_view uri_tail(std::string_view s) {
  const size_t p = s.rfind('/');
  return p == std::string_view::npos ? s : s.substr(p + 1);
}

extern "C" int kemr_idmap_create(const char* blob, const int64_t* offsets, int64_t n, kemr_idmap_t** out) {
  if (!offsets || !out || n < 0 || (n && !blob)) return fail(KEMR_ERR_ARG, "idmap_create: bad argument");
  kemr_idmap* m = new kemr_idmap();
  m->blob.assign(blob ? blob : "", (size_t)offsets[n]);
  m->map.reserve((size_t)n * 2);
  for (int64_t i = 0; i < n; ++i) {
    if (offsets[i + 1] < offsets[i]) { delete m; return fail(KEMR_ERR_ARG, "idmap_create: offsets must not decrease"); }
    // a repeated uuid keeps its LAST row, like the reference's {uuid: idx for idx, uuid in enumerate(...)} (fusion.py:62)
    m->map[std::string_view(m->blob.data() + offsets[i], (size_t)(offsets[i + 1] - offsets[i]))] = i;
  }
  *out = m;
  return KEMR_OK;
}

extern "C" int kemr_idmap_destroy(kemr_idmap_t* m) { delete m; return KEMR_OK; }

extern "C" int kemr_idmap_lookup(const kemr_idmap_t* m, const char* blob, const int64_t* offsets, int64_t n,
                                 int normalize_uri, int64_t* out_rows) {
  if (!m || !offsets || !out_rows || n < 0 || (n && !blob)) return fail(KEMR_ERR_ARG, "idmap_lookup: bad argument");
  for (int64_t i = 0; i < n; ++i) {
    std::string_view key(blob + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
    if (normalize_uri) key = uri_tail(key);
    auto it = m->map.find(key);
    out_rows[i] = it == m->map.end() ? -1 : it->second;
  }
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- persisted bf16 embedding store
// File = 64-byte header + M*D bf16 values, row-major.  Replaces the reference's `data/embeddings` directory
// (clip_retrieval.py:28,35; contents defined by remote code) with a layout a shard can be cut from by byte range.
struct StoreHeader {
  char magic[8];          // "KEMRSTOR"
  uint32_t version;       // 1
  uint32_t dtype;         // 1 = bf16
  int64_t rows;
  int32_t dim;
  int32_t reserved[9];
};
static_assert(sizeof(StoreHeader) == 64, "store header is 64 bytes");

static int store_read_header(int fd, const char* path, StoreHeader* h) {
  if (pread(fd, h, sizeof *h, 0) != (ssize_t)sizeof *h) return fail(KEMR_ERR_ARG, "store: %s is shorter than a header", path);
  if (memcmp(h->magic, "KEMRSTOR", 8) != 0 || h->version != 1 || h->dtype != 1 || h->rows < 0 || h->dim <= 0)
    return fail(KEMR_ERR_ARG, "store: %s is not a version-1 bf16 embedding store", path);
  return KEMR_OK;
}

extern "C" int kemr_store_write(const char* path, const uint16_t* rows_host, int64_t M, int D) {
  if (!path || (M && !rows_host) || M < 0 || D <= 0 || D % 8) return fail(KEMR_ERR_ARG, "store_write: bad argument");
  const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return fail(KEMR_ERR_ARG, "store_write: cannot create %s", path);
  StoreHeader h{};
  memcpy(h.magic, "KEMRSTOR", 8); h.version = 1; h.dtype = 1; h.rows = M; h.dim = D;
  bool ok = write(fd, &h, sizeof h) == (ssize_t)sizeof h;
  const char* p = reinterpret_cast<const char*>(rows_host);
  size_t left = (size_t)M * D * 2;
  while (ok && left) {
    const ssize_t w = write(fd, p, std::min<size_t>(left, (size_t)1 << 30));
    if (w <= 0) ok = false; else { p += w; left -= (size_t)w; }
  }
  close(fd);
  return ok ? KEMR_OK : fail(KEMR_ERR_ARG, "store_write: short write to %s", path);
}

extern "C" int kemr_store_info(const char* path, int64_t* rows, int* dim) {
  if (!path) return fail(KEMR_ERR_ARG, "store_info: null path");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(KEMR_ERR_ARG, "store_info: cannot open %s", path);
  StoreHeader h;
  int rc = store_read_header(fd, path, &h);
  struct stat stt;
  if (!rc && (fstat(fd, &stt) != 0 || (int64_t)stt.st_size < (int64_t)sizeof h + h.rows * h.dim * 2))
    rc = fail(KEMR_ERR_ARG, "store_info: %s is truncated", path);
  close(fd);
  if (rc) return rc;
  if (rows) *rows = h.rows;
  if (dim) *dim = h.dim;
  return KEMR_OK;
}

// rows [row_lo, row_hi) -> device memory: pread into two page-locked buffers, each copy in flight while the next
// chunk is read.  Synchronises `stream` before returning (the staging buffers are freed).
extern "C" int kemr_store_load(const char* path, int64_t row_lo, int64_t row_hi, uint16_t* dst_device, kemr_stream_t stream) {
  if (!path || !dst_device || row_lo < 0 || row_hi < row_lo) return fail(KEMR_ERR_ARG, "store_load: bad argument");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(KEMR_ERR_ARG, "store_load: cannot open %s", path);
  StoreHeader h;
  int rc = store_read_header(fd, path, &h);
  if (!rc && row_hi > h.rows) rc = fail(KEMR_ERR_ARG, "store_load: rows [%lld, %lld) beyond the %lld rows of %s", (long long)row_lo, (long long)row_hi, (long long)h.rows, path);
  if (rc) { close(fd); return rc; }
  const size_t row_bytes = (size_t)h.dim * 2, total = (size_t)(row_hi - row_lo) * row_bytes;
  const size_t chunk = std::min<size_t>(std::max<size_t>(total, 1), (size_t)32 << 20);
  char* stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t st = S(stream);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMallocHost(&stage[i], chunk);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
  }
  size_t off = 0;
  int b = 0;
  bool io_ok = true;
  while (e == cudaSuccess && io_ok && off < total) {
    const size_t n = std::min(chunk, total - off);
    e = cudaEventSynchronize(done[b]);                      // the copy that last used this buffer has finished
    size_t got = 0;
    while (e == cudaSuccess && got < n) {
      const ssize_t r = pread(fd, stage[b] + got, n - got, (off_t)(sizeof h + (size_t)row_lo * row_bytes + off + got));
      if (r <= 0) { io_ok = false; break; }
      got += (size_t)r;
    }
    if (e == cudaSuccess && io_ok) e = cudaMemcpyAsync(reinterpret_cast<char*>(dst_device) + off, stage[b], n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && io_ok) e = cudaEventRecord(done[b], st);
    off += n;
    b ^= 1;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  for (int i = 0; i < 2; ++i) { if (done[i]) cudaEventDestroy(done[i]); if (stage[i]) cudaFreeHost(stage[i]); }
  close(fd);
  if (!io_ok) return fail(KEMR_ERR_ARG, "store_load: short read from %s", path);
  if (e != cudaSuccess) return fail(KEMR_ERR_CUDA, "store_load: %s", cudaGetErrorString(e));
  return KEMR_OK;
}

// %globaltimer (ns) at the phase boundaries of the selection of query 0 in the most recent launch: entry, lists merged
// (A1), candidates cut (A2), pruned (A3), KG hits joined (B), re-scored (C), ordered and written (D).  Debug build
// only (libkemr_debug.so); the release library reports zeros.
extern "C" int kemr_debug_select_stamps(int64_t* out16_host) {
  if (!out16_host) return fail(KEMR_ERR_ARG, "debug_select_stamps: null pointer");
  memset(out16_host, 0, 16 * 8);
#ifdef KEMR_DEBUG
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out16_host, g_sel_stamps, 16 * 8));
#endif
  return KEMR_OK;
}

// ----------------------------------------------------------------------------- plan introspection (no device needed)
// The tcgen05 plan for a shape on a hypothetical device with `sms` SMs and room for `quads` clusters of four:
// out[0..15] = parts, q_pad, n_tile, n_qb, n_t, ctas, stages, kc, K, cl, gran, vq, all_slots, two, merged, q_blk;
// returns KEMR_ERR_UNSUPPORTED when the shape cannot be planned.  Lets the host-side scheduling logic (unit ranges,
// part slots) be property-tested on a CPU (tests/test_plan_cpu.py).
extern "C" int kemr_debug_mma_plan(int Q, int64_t M, int D, int galleries, int k_sel, int equal_weights, int sms, int quads,
                                   int64_t* out16) {
  if (!out16 || Q <= 0 || M <= 0 || galleries < 1 || galleries > 2) return fail(KEMR_ERR_ARG, "debug_mma_plan: bad argument");
  if (!mma_supported(D, k_sel)) return fail(KEMR_ERR_UNSUPPORTED, "debug_mma_plan: unsupported D / k_sel");
  MmaPlan p;
  if (mma_make_plan(Q, M, D, galleries, k_sel, kModeTopk, sms, quads, equal_weights != 0, &p))
    return fail(KEMR_ERR_UNSUPPORTED, "debug_mma_plan: shape cannot be planned");
  const int64_t v[16] = {p.parts, p.q_pad, p.n_tile, p.n_qb, p.n_t, p.ctas, p.stages, p.kc, p.K, p.cl, p.gran, p.vq,
                         p.all_slots, p.two, p.merged, p.q_blk};
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  return KEMR_OK;
}
