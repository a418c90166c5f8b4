// Shared device/host helpers for libkemr (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace kemr {

constexpr int kWarp = 32;
constexpr int kMaxKSel = 128;          // candidates kept per query by the fp32 scan
constexpr int kMaxD = 1024;

// ----------------------------------------------------------------------------- bf16 helpers
__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }
__device__ __forceinline__ float bf16_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

__device__ __forceinline__ uint16_t f32_to_bf16_rne(float x) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);   // quiet NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ----------------------------------------------------------------------------- ordering keys
// A candidate is a 64-bit key: high word = order-preserving image of the fp32 score, low word =
// ~row, so that "larger key" == "higher score, then lower row index".  0 is the empty slot.
__host__ __device__ __forceinline__ uint32_t order_f32(float f) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
  if ((u & 0x7fffffffu) > 0x7f800000u) return 1u;          // NaN: below -inf, above "empty"
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unorder_f32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  if (k == 1u) u = 0x7fc00000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return ((uint64_t)order_f32(score) << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return unorder_f32((uint32_t)(k >> 32)); }

// ----------------------------------------------------------------------------- warp primitives
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact bf16 -> binary64 (bf16 -> fp32 is a shift, fp32 -> fp64 is exact incl. subnormals)
__device__ __forceinline__ double bf16_to_f64(uint32_t b16) { return (double)__uint_as_float(b16 << 16); }
__device__ __forceinline__ double bf16hi_to_f64(uint32_t packed) { return (double)__uint_as_float(packed & 0xffff0000u); }

// Canonical binary64 dot product (oracle/oracle.py::canon_dot64): the row is cut into 16-byte
// pieces of 8 bf16; lane l owns pieces l, l+32, ... (exactly what it fetches with coalesced
// 16-byte loads) and keeps one running sum per element position of its pieces (products of two
// bf16 are exact in binary64), combines the 8 sums with a fixed tree, and the 32 lanes then fold
// 16,8,4,2,1.  NP = pieces per lane = ceil(D/256).
constexpr int kCanonPieces = 4;

// raw 16-byte pieces of one row owned by this lane
template <int NP>
struct CanonRow {
  uint4 x[NP];
  __device__ __forceinline__ void load(const uint16_t* __restrict__ g, int D, int lane) {
    const int npiece = D >> 3;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int c = lane + 32 * j;
      x[j] = (c < npiece) ? ldg_stream(g + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
    }
  }
};

// the lane's share of a query row, widened to binary64 once and reused for many gallery rows
template <int NP>
struct CanonQueryT {
  double v[NP][8];
  __device__ __forceinline__ void load(const uint16_t* __restrict__ q, int D, int lane) {
    const int npiece = D >> 3;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int c = lane + 32 * j;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (c < npiece) x = *reinterpret_cast<const uint4*>(q + (size_t)c * 8);
      v[j][0] = bf16_to_f64(x.x & 0xffffu); v[j][1] = bf16hi_to_f64(x.x);
      v[j][2] = bf16_to_f64(x.y & 0xffffu); v[j][3] = bf16hi_to_f64(x.y);
      v[j][4] = bf16_to_f64(x.z & 0xffffu); v[j][5] = bf16hi_to_f64(x.z);
      v[j][6] = bf16_to_f64(x.w & 0xffffu); v[j][7] = bf16hi_to_f64(x.w);
    }
  }
  // every lane returns the canonical dot with an already-fetched row: 8 independent running sums
  // per lane (one per element of its 16-byte pieces, pieces in increasing order), an 8-to-1 tree,
  // then the 32 lanes fold 16,8,4,2,1
  __device__ __forceinline__ double dot(const CanonRow<NP>& r, int D, int lane) const {
    const int npiece = D >> 3;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0, p7 = 0.0;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (lane + 32 * j < npiece) {
        const uint4 x = r.x[j];
        p0 = fma(v[j][0], bf16_to_f64(x.x & 0xffffu), p0); p1 = fma(v[j][1], bf16hi_to_f64(x.x), p1);
        p2 = fma(v[j][2], bf16_to_f64(x.y & 0xffffu), p2); p3 = fma(v[j][3], bf16hi_to_f64(x.y), p3);
        p4 = fma(v[j][4], bf16_to_f64(x.z & 0xffffu), p4); p5 = fma(v[j][5], bf16hi_to_f64(x.z), p5);
        p6 = fma(v[j][6], bf16_to_f64(x.w & 0xffffu), p6); p7 = fma(v[j][7], bf16hi_to_f64(x.w), p7);
      }
    }
    double acc = __dadd_rn(__dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3)), __dadd_rn(__dadd_rn(p4, p5), __dadd_rn(p6, p7)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, o));
    return __shfl_sync(0xffffffffu, acc, 0);
  }
  // canonical dots with TWO already-fetched rows (T2I and T2T row of one candidate), the two reduction
  // chains interleaved; same operation order per row as dot().  Valid on lane 0 only.
  __device__ __forceinline__ void dot2_lane0(const CanonRow<NP>& ra, const CanonRow<NP>& rb, int D, int lane,
                                             double& out_a, double& out_b) const {
    const int npiece = D >> 3;
    double a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = 0.0; b[e] = 0.0; }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (lane + 32 * j < npiece) {
        const uint4 x = ra.x[j], y = rb.x[j];
        a[0] = fma(v[j][0], bf16_to_f64(x.x & 0xffffu), a[0]); b[0] = fma(v[j][0], bf16_to_f64(y.x & 0xffffu), b[0]);
        a[1] = fma(v[j][1], bf16hi_to_f64(x.x), a[1]);         b[1] = fma(v[j][1], bf16hi_to_f64(y.x), b[1]);
        a[2] = fma(v[j][2], bf16_to_f64(x.y & 0xffffu), a[2]); b[2] = fma(v[j][2], bf16_to_f64(y.y & 0xffffu), b[2]);
        a[3] = fma(v[j][3], bf16hi_to_f64(x.y), a[3]);         b[3] = fma(v[j][3], bf16hi_to_f64(y.y), b[3]);
        a[4] = fma(v[j][4], bf16_to_f64(x.z & 0xffffu), a[4]); b[4] = fma(v[j][4], bf16_to_f64(y.z & 0xffffu), b[4]);
        a[5] = fma(v[j][5], bf16hi_to_f64(x.z), a[5]);         b[5] = fma(v[j][5], bf16hi_to_f64(y.z), b[5]);
        a[6] = fma(v[j][6], bf16_to_f64(x.w & 0xffffu), a[6]); b[6] = fma(v[j][6], bf16_to_f64(y.w & 0xffffu), b[6]);
        a[7] = fma(v[j][7], bf16hi_to_f64(x.w), a[7]);         b[7] = fma(v[j][7], bf16hi_to_f64(y.w), b[7]);
      }
    }
    double sa = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])), __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
    double sb = __dadd_rn(__dadd_rn(__dadd_rn(b[0], b[1]), __dadd_rn(b[2], b[3])), __dadd_rn(__dadd_rn(b[4], b[5]), __dadd_rn(b[6], b[7])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ta = __shfl_down_sync(0xffffffffu, sa, o), tb = __shfl_down_sync(0xffffffffu, sb, o);
      sa = __dadd_rn(sa, ta);
      sb = __dadd_rn(sb, tb);
    }
    out_a = sa; out_b = sb;
  }
};
using CanonQuery = CanonQueryT<kCanonPieces>;

__device__ __forceinline__ double canon_dot_q(const CanonQuery& cq, const uint16_t* __restrict__ g, int D, int lane) {
  CanonRow<kCanonPieces> r;
  r.load(g, D, lane);
  return cq.dot(r, D, lane);
}

// convenience: one-off pair
__device__ __forceinline__ double canon_dot_warp(const uint16_t* __restrict__ a,
                                                 const uint16_t* __restrict__ b, int D, int lane) {
  CanonQuery cq;
  cq.load(a, D, lane);
  return canon_dot_q(cq, b, D, lane);
}

// final = fl(fl(alpha * fl(fl(w_a*S_a) + fl(w_b*S_b))) + bonus); no contraction allowed.
__device__ __forceinline__ double canon_fuse(double sa, double sb, bool two, double wa, double wb,
                                             double alpha, double bonus, bool has_bonus) {
  double clip = __dmul_rn(wa, sa);
  if (two) clip = __dadd_rn(clip, __dmul_rn(wb, sb));
  double f = __dmul_rn(alpha, clip);
  if (has_bonus) f = __dadd_rn(f, bonus);
  return f;
}

// "a ranks ahead of b": higher score first, then lower index; NaN after everything.
__device__ __forceinline__ bool ahead64(double sa, int64_t ia, double sb, int64_t ib) {
  const bool na = isnan(sa), nb = isnan(sb);
  if (na || nb) return (!na && nb) || (na && nb && ia < ib);
  return sa > sb || (sa == sb && ia < ib);
}

// Insert key x into a descending sorted list L[0..K) held in shared memory, cooperatively by one
// warp.  Precondition: x > L[K-1] (caller checked the threshold).  K <= 128.
__device__ __forceinline__ void warp_list_insert(uint64_t* L, int K, uint64_t x, int lane) {
  uint64_t nv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = lane + 32 * i;
    if (p < K) {
      const uint64_t cur = L[p];
      const uint64_t prev = p > 0 ? L[p - 1] : ~0ull;
      nv[i] = cur > x ? cur : (prev > x ? x : prev);
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = lane + 32 * i;
    if (p < K) L[p] = nv[i];
  }
  __syncwarp();
}

}  // namespace kemr
