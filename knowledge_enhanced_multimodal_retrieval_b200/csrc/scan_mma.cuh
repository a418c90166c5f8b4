// TMA-fed tcgen05 / TMEM similarity scan with fused fusion-weight + top-k / count epilogue
// (north_star kernels 1a + 2).  sm_100a only; hand-written PTX, no CUTLASS.
//
// Orientation: D[128 x N] (fp32, TMEM) = Qblk[128 x D] * Gtile[N x D]^T, both operands bf16,
// K-major, staged in shared memory by TMA with the 128-byte swizzle:
//   A = 128 query rows (rows past the batch are never loaded; their accumulator lanes are ignored)
//   B = N_TILE gallery rows (TMA zero-fills rows past the end of the shard and columns past D)
// so one TMEM lane == one query and the epilogue thread that owns the lane walks the gallery
// columns.  With two galleries (T2I + T2T) each gets its own accumulator and the epilogue forms
// w_a*acc_a + w_b*acc_b in fp32.  Accumulators are double buffered in TMEM (2 x G x N_TILE <= 512
// columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation),
// warps 2..5 = epilogue (one TMEM lane quadrant each).
// Work: the (query block, gallery tile) grid is flattened and cut into equal contiguous ranges, one
// per persistent CTA (<= #SMs); a CTA emits one K-entry candidate list per query for every query
// block its range touches ("part slot"), merged later by select.cuh.
//
// Epilogue per element: 1-2 FFMA for the fusion weights, one compare against the thread's running
// threshold; the top-K list of a query lives in REGISTERS (sorted, K <= 32) and an insert is a
// branch-free compare/select chain, so the score matrix never exists outside TMEM.
// Roofline: tensor pipe for batch >= ~250 (2*B*G*M*D flop), HBM below (G*M*D*2 bytes).
#pragma once
#include <cuda.h>
#include "scan_warp.cuh"

namespace kemr {

constexpr int kMmaThreads = 192;
constexpr int kBlockM = 128;           // queries per block == TMEM lanes
constexpr int kBlockK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kSmemBudget = 227 * 1024;

struct MmaPlan {
  int parts = 0;          // part slots per query block
  int q_pad = 0;          // queries rounded up to kBlockM
  int n_tile = 0;         // gallery rows per MMA tile (128 with two galleries, else 256)
  int n_qb = 0, n_t = 0;  // query blocks, gallery tiles
  int ctas = 0;
  int stages = 0;
  int kc = 0;             // K chunks of 64
  int a_rows = 0;         // query rows actually loaded per block
  int K = 0;              // list length (8, 16, 24, 32)
  size_t smem = 0;
};

inline bool mma_built() { return true; }
inline bool mma_supported(int D, int K) { return D % 8 == 0 && D >= 8 && D <= kMaxD && K >= 1 && K <= 32; }

static thread_local char g_mma_error[256] = "";
inline const char* mma_last_error() { return g_mma_error; }

inline int mma_round_k(int K) { return K <= 8 ? 8 : K <= 16 ? 16 : K <= 24 ? 24 : 32; }

inline int mma_make_plan(int Q, int64_t M, int D, int G, int K, int mode, int sms, MmaPlan* p) {
  (void)mode;
  p->n_tile = G == 2 ? 128 : 256;
  p->n_qb = (Q + kBlockM - 1) / kBlockM;
  p->q_pad = p->n_qb * kBlockM;
  const int64_t nt = (M + p->n_tile - 1) / p->n_tile;
  if (nt > (1ll << 30)) return 1;
  p->n_t = (int)nt;
  const int64_t W = (int64_t)p->n_qb * p->n_t;
  p->ctas = (int)std::min<int64_t>(sms, W);
  // widest span of CTAs touching one query block
  int parts = 1;
  for (int qb = 0; qb < p->n_qb; ++qb) {
    const int64_t w0 = (int64_t)qb * p->n_t, w1 = w0 + p->n_t - 1;
    const int c0 = (int)(((w0 + 1) * p->ctas - 1) / W), c1 = (int)(((w1 + 1) * p->ctas - 1) / W);
    parts = std::max(parts, c1 - c0 + 1);
  }
  p->parts = parts;
  p->kc = (D + kBlockK - 1) / kBlockK;
  p->a_rows = Q >= kBlockM ? kBlockM : (Q + 7) / 8 * 8;
  p->K = mma_round_k(K);
  const size_t stage = (size_t)kBlockM * 128 + (size_t)p->n_tile * 128;
  p->stages = (int)std::min<size_t>(8, (kSmemBudget - 2048) / stage);
  p->smem = (size_t)p->stages * stage + 2048;
  return p->stages >= 2 ? 0 : 1;
}

// ------------------------------------------------------------------------------------ PTX wrappers
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct MmaArgs {
  ScanArgs s;
  int n_tile, n_qb, n_t, stages, kc, a_rows, parts, q_pad;
  long long W;
};

// sorted (descending) per-thread candidate list in registers; rows arrive in increasing order, so
// strict '>' keeps the lower index ahead among equal scores
template <int K>
struct RegList {
  float sc[K];
  uint32_t ix[K];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int p = 0; p < K; ++p) { sc[p] = -INFINITY; ix[p] = 0xffffffffu; }
  }
  __device__ __forceinline__ float threshold() const { return sc[K - 1]; }
  __device__ __forceinline__ void insert(float s, uint32_t r) {
    bool up_next = true;                       // s > sc[K-1] is the caller's precondition
#pragma unroll
    for (int p = K - 1; p >= 1; --p) {
      const bool up = s > sc[p - 1];           // new entry passes position p-1 -> p-1 moves down to p
      const float nsc = up ? sc[p - 1] : (up_next ? s : sc[p]);
      const uint32_t nix = up ? ix[p - 1] : (up_next ? r : ix[p]);
      sc[p] = nsc; ix[p] = nix;
      up_next = up;
    }
    if (up_next) { sc[0] = s; ix[0] = r; }
  }
};

template <int K>
__global__ void __launch_bounds__(kMmaThreads, 1)
scan_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g0,
                const __grid_constant__ CUtensorMap map_g1, MmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int n_tile = a.n_tile;
  const uint32_t a_bytes = kBlockM * 128, b_bytes = (uint32_t)n_tile * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  unsigned char* bar_base = smem + (size_t)a.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.s.G;
  const long long w_lo = a.W * blockIdx.x / gridDim.x, w_hi = a.W * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 4); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&map_q); ptx::prefetch_tmap(&map_g0);
    if (G > 1) ptx::prefetch_tmap(&map_g1);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)a.a_rows * 128 + b_bytes;
      for (long long w = w_lo; w < w_hi; ++w) {
        const int qb = (int)(w / a.n_t), t = (int)(w % a.n_t);
        for (int g = 0; g < G; ++g) {
          const CUtensorMap* mg = g ? &map_g1 : &map_g0;
          for (int kc = 0; kc < a.kc; ++kc) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            unsigned char* sa = smem + (size_t)stage * stage_bytes;
            ptx::mbar_expect_tx(&full_bar[stage], tx);
            ptx::tma_load_2d(sa, &map_q, &full_bar[stage], kc * kBlockK, qb * kBlockM);
            ptx::tma_load_2d(sa + a_bytes, mg, &full_bar[stage], kc * kBlockK, t * n_tile);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t idesc = umma_idesc_bf16(kBlockM, n_tile);
      long long it = 0;
      for (long long w = w_lo; w < w_hi; ++w, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t bphase = (uint32_t)((it >> 1) & 1);
        ptx::mbar_wait(&tempty_bar[buf], bphase ^ 1);
        ptx::tc_fence_after();
        for (int g = 0; g < G; ++g) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * G + g) * (uint32_t)n_tile;
          for (int kc = 0; kc < a.kc; ++kc) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + a_bytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              ptx::mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
            ptx::mma_commit(&empty_bar[stage]);          // frees the smem stage when these MMAs retire
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
        ptx::mma_commit(&tfull_bar[buf]);                // accumulators of this tile are complete
      }
    }
  } else {
    // ================================================================= epilogue (warps 2..5)
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int qrow = quad * 32 + lane;                   // query row inside the block
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float w0 = a.s.w[0], w1 = a.s.w[1];
    const int mode = a.s.mode;
    RegList<K> list;
    list.reset();
    int32_t cnt = 0;
    float blo = 0.f, bhi = 0.f;
    int cur_qb = -1;
    const long long Wt = a.W;
    const int C = gridDim.x;
    long long it = 0;
    auto flush = [&](int qb) {
      if (qb < 0) return;
      const long long wq = (long long)qb * a.n_t;
      const int c_first = (int)(((wq + 1) * C - 1) / Wt);
      const int slot = (int)blockIdx.x - c_first;
      const int qg = qb * kBlockM + qrow;
      if (mode == kModeTopk) {
        uint64_t* dst = a.s.part_keys + ((size_t)slot * a.q_pad + qg) * a.s.K;
#pragma unroll
        for (int p = 0; p < K; ++p)
          if (p < a.s.K) dst[p] = list.ix[p] == 0xffffffffu ? 0ull : make_key(list.sc[p], list.ix[p]);
      } else if (mode == kModeCount) {
        a.s.part_count[(size_t)slot * a.q_pad + qg] = cnt;
      }
    };
    for (long long w = w_lo; w < w_hi; ++w, ++it) {
      const int qb = (int)(w / a.n_t), t = (int)(w % a.n_t);
      if (qb != cur_qb) {
        flush(cur_qb);
        cur_qb = qb;
        list.reset();
        cnt = 0;
        if (mode == kModeCount) {
          const int qg = qb * kBlockM + qrow;
          blo = a.s.band_lo[qg]; bhi = a.s.band_hi[qg];   // padded rows hold +huge: never count
        }
      }
      const int buf = (int)(it & 1);
      const uint32_t bphase = (uint32_t)((it >> 1) & 1);
      ptx::mbar_wait(&tfull_bar[buf], bphase);
      ptx::tc_fence_after();
      const int qg = qb * kBlockM + qrow;
      const bool qvalid = qg < a.s.Q;
      const long long row0 = (long long)t * n_tile;
      const int ncols = (int)min((long long)n_tile, a.s.M - row0);
      float thr = list.threshold();
      for (int c0 = 0; c0 < n_tile; c0 += 32) {
        if (c0 >= ncols) break;                          // warp-uniform
        uint32_t ra[32], rb[32];
        ptx::tmem_ld32(lane_addr + (uint32_t)(buf * G) * (uint32_t)n_tile + (uint32_t)c0, ra);
        if (G > 1) ptx::tmem_ld32(lane_addr + (uint32_t)(buf * G + 1) * (uint32_t)n_tile + (uint32_t)c0, rb);
        ptx::tmem_ld_wait();
        if (mode == kModeTopk) {
          uint32_t mask = 0;
          float sv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float s = w0 * __uint_as_float(ra[j]);
            if (G > 1) s = fmaf(w1, __uint_as_float(rb[j]), s);
            sv[j] = s;
            mask |= (s > thr && c0 + j < ncols) ? (1u << j) : 0u;
          }
          if (!qvalid) mask = 0;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            // dynamic register index: resolve with a select chain over the unrolled chunk
            float s = sv[0];
#pragma unroll
            for (int jj = 1; jj < 32; ++jj) s = (jj == j) ? sv[jj] : s;
            if (s > thr) {
              list.insert(s, (uint32_t)(row0 + c0 + j));
              thr = list.threshold();
            }
          }
        } else if (mode == kModeCount) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float s = w0 * __uint_as_float(ra[j]);
            if (G > 1) s = fmaf(w1, __uint_as_float(rb[j]), s);
            const bool in = c0 + j < ncols;
            if (in && s > bhi) ++cnt;
            else if (in && s >= blo) {
              const unsigned int slot = atomicAdd(a.s.amb_counter, 1u);
              if (slot < a.s.amb_cap) { a.s.amb_q[slot] = (uint32_t)qg; a.s.amb_row[slot] = (uint32_t)(row0 + c0 + j); }
            }
          }
        } else {
          if (qvalid) {
            float* dst = a.s.dense + (size_t)qg * a.s.ld + row0 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float s = w0 * __uint_as_float(ra[j]);
              if (G > 1) s = fmaf(w1, __uint_as_float(rb[j]), s);
              if (c0 + j < ncols) dst[j] = s;
            }
          }
        }
      }
      // release this accumulator buffer to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
    }
    flush(cur_qb);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

inline int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { snprintf(g_mma_error, sizeof g_mma_error, "cuTensorMapEncodeTiled unavailable"); return 1; }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_mma_error, sizeof g_mma_error, "cuTensorMapEncodeTiled failed (%d) rows=%lld D=%d box=%d", (int)r,
             (long long)rows, D, box_rows);
    return 1;
  }
  return 0;
}

template <int K>
inline int mma_launch_k(const CUtensorMap& mq, const CUtensorMap& m0, const CUtensorMap& m1, const MmaArgs& ma,
                        const MmaPlan& pl, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(scan_mma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem + 1024);
  if (e != cudaSuccess) { snprintf(g_mma_error, sizeof g_mma_error, "smem attribute: %s", cudaGetErrorString(e)); return 1; }
  scan_mma_kernel<K><<<pl.ctas, kMmaThreads, pl.smem + 1024, st>>>(mq, m0, m1, ma);
  e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_mma_error, sizeof g_mma_error, "launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

inline int mma_launch(const ScanArgs& s, const MmaPlan& pl, cudaStream_t st) {
  CUtensorMap mq, m0, m1;
  if (make_tmap_2d(&mq, s.q, s.Q, s.D, pl.a_rows)) return 1;
  if (make_tmap_2d(&m0, s.gal[0], s.M, s.D, pl.n_tile)) return 1;
  if (s.G > 1) { if (make_tmap_2d(&m1, s.gal[1], s.M, s.D, pl.n_tile)) return 1; }
  else m1 = m0;
  MmaArgs ma;
  ma.s = s;
  ma.n_tile = pl.n_tile; ma.n_qb = pl.n_qb; ma.n_t = pl.n_t; ma.stages = pl.stages; ma.kc = pl.kc;
  ma.a_rows = pl.a_rows; ma.parts = pl.parts; ma.q_pad = pl.q_pad;
  ma.W = (long long)pl.n_qb * pl.n_t;
  switch (pl.K) {
    case 8: return mma_launch_k<8>(mq, m0, m1, ma, pl, st);
    case 16: return mma_launch_k<16>(mq, m0, m1, ma, pl, st);
    case 24: return mma_launch_k<24>(mq, m0, m1, ma, pl, st);
    default: return mma_launch_k<32>(mq, m0, m1, ma, pl, st);
  }
}

}  // namespace kemr
