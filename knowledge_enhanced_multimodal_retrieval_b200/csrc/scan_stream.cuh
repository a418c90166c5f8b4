// HBM-streaming warp-dot search for small, latency-bound query batches, scan + selection in ONE launch
// (north_star kernel 1b + 2 for batch 1..4; BASELINE config 3).
//
// One CTA of sixteen warps per SM; a CTA owns a contiguous range of gallery rows.  A warp takes two rows per step:
// fully coalesced 16-byte loads that bypass L1 (lane l owns pieces l, l+32, ...), all issued before the first use, so
// sixteen warps keep ~100 KB per SM in flight -- for a 132 MB gallery (43 000 x 768-d x 2) a plain read kernel of
// this shape is the fastest way through the data (tools/microbench/stream_floor.cu: 28.8 us, against 32.8 us for a
// bulk-copy / mbarrier ring, which only wins on multi-GB scans where its set-up is amortised).  fp32 FMAs in
// independent chains against the queries held in registers (fusion weights applied to the per-lane partial sums, so
// T2I + T2T cost one reduction), ONE transposed shuffle reduction for all (row, query) sums of the warp (2..8 values:
// 5..9 shuffles instead of 5 per value), and the lanes that end up holding a sum append (score, row) keys that beat
// the CTA's running threshold to a candidate buffer in shared memory.  Nothing is sorted while rows stream: the
// buffer is cut back to its K best by counting only when it fills (never, for a 43 k-row gallery) and once at the
// end, which leaves one K-entry list per CTA and query.
// The last CTA to finish (device-wide arrival counter) then runs the selection stage of select.cuh in the same
// launch: merge of the lists, KG hits, canonical binary64 re-scoring, final order, certificate.  No second launch
// (an empty launch costs ~6 us event to event on this part), no score matrix, no host round trip.
// Algorithmic traffic: G*M*D*2 bytes per group of QB queries (roofline: MEASURED_PEAKS hbm_gbs).
#pragma once
#include "scan_mma.cuh"     // ptx:: helpers
#include "select.cuh"

namespace kemr {

constexpr int kStreamWarps = 16;                          // one 512-thread CTA per SM: half as many lists to merge as 2 x 8 warps, twice the threads in the selection
constexpr int kStreamThreads = kStreamWarps * 32;
constexpr int kStreamRW = 2;                              // rows per warp and step
constexpr int kStreamRows = kStreamWarps * kStreamRW;     // rows per CTA and step
constexpr int kStreamCand = 1024;                         // candidate keys per query a CTA can hold between cuts
constexpr int kStreamCheck = 16;                          // steps between two looks at the buffer fill (<= 256 appends)

struct StreamArgs {
  ScanArgs s;               // queries, galleries, weights; s.K = list length per CTA; s.part_keys [P][Q][K]
  SelectArgs sel;           // the selection stage (P = gridDim.x lists per query)
  unsigned int* done;       // [groups] arrival counters, zero on entry, reset by the last CTA
  long long* stamps;        // optional [3] globaltimer stamps of group 0: first CTA start, last scan arrival, select end
  unsigned int cand_off;    // offset of the candidate buffers behind the selection scratch
  // optional: queries arrive as fp32 [Q][D] (device memory) and are rounded to bf16 in this kernel -- optionally
  // L2-normalised first -- with exactly the arithmetic of quantize_rows_kernel; CTA 0 of a group stores the bf16 rows
  // to s.q for the selection stage.  Saves the separate quantise launch of the host-buffer search.
  const float* q_f32;
  int q_normalize;
};

inline size_t stream_cand_off(size_t select_bytes) { return (select_bytes + 15) / 16 * 16; }
// layout behind the selection scratch: cand[QB][kStreamCand] u64 | best[QB][kMaxKSel] u64 | bf16 queries [QB][kMaxD]
inline size_t stream_smem_bytes(int QB, size_t select_bytes) {
  return stream_cand_off(select_bytes) + (size_t)QB * kStreamCand * 8 + (size_t)QB * kMaxKSel * 8 + (size_t)QB * kMaxD * 2 + 64;
}

namespace ptx {
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
}  // namespace ptx

// keep the K best of cand[0..n) (distinct keys): ranks by counting, all threads of the CTA
__device__ __forceinline__ void stream_cut(const uint64_t* cand, uint64_t* best, int n, int K) {
  for (int i = threadIdx.x; i < n; i += kStreamThreads) {
    const uint64_t x = cand[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += cand[j] > x ? 1 : 0;
    if (r < K) best[r] = x;
  }
}

// Sum N per-lane values over the warp at once: every step halves the payload (lanes with the step's bit set keep the
// upper half) until one value per lane is left, then plain butterfly steps.  Lane l returns the total of value
// stream_value_of<N>(l); the lanes whose low bits differ hold copies.
template <int N>
__device__ __forceinline__ float stream_reduce(float (&v)[N], int lane) {
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int step = 16 >> j;
    const int nn = N >> (j + 1);                             // values kept after this step (0: already down to one)
    if (nn >= 1) {
      const bool up = (lane & step) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < nn) {
          const float send = up ? v[i] : v[i + nn];
          const float keep = up ? v[i + nn] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], step);
    }
  }
  return v[0];
}
template <int N>
__device__ __forceinline__ int stream_value_of(int lane) {
  int idx = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int nn = N >> (j + 1);
    if (nn >= 1 && (lane & (16 >> j))) idx += nn;
  }
  return idx;
}

template <int QB, int CH, int NP>
__global__ void __launch_bounds__(kStreamThreads, 1) scan_stream_kernel(const __grid_constant__ StreamArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem + a.cand_off);     // [QB][kStreamCand]
  uint64_t* best = cand + (size_t)QB * kStreamCand;                     // [QB][kMaxKSel]
  uint16_t* qs = reinterpret_cast<uint16_t*>(best + (size_t)QB * kMaxKSel);   // [QB][D] bf16 queries (fp32-query mode)
  __shared__ unsigned int s_cnt[QB];
  __shared__ unsigned long long s_thr[QB];
  __shared__ int s_last;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = blockIdx.y;
  const int q0 = group * QB;
  const int nq = min(QB, a.s.Q - q0);
  const int P = gridDim.x;
  const int64_t r0 = (a.s.M * (int64_t)blockIdx.x) / P;
  const int64_t r1 = (a.s.M * (int64_t)(blockIdx.x + 1)) / P;
  const int niter = (int)((r1 - r0 + kStreamRows - 1) / kStreamRows);
  const int K = a.s.K;
  const int D = a.s.D;
  const int nchunk = D >> 3;

  if (a.stamps && group == 0 && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(a.stamps), ptx::globaltimer());
  if (threadIdx.x < QB) { s_cnt[threadIdx.x] = 0; s_thr[threadIdx.x] = 0; }

  // queries -> fp32 registers (lane l keeps the d-slices it will meet in every row)
  const float* q_f32 = a.q_f32;
  const int64_t* hit_rowptr = a.sel.hit_rowptr;
  const int32_t* hit_col = a.sel.hit_col;
  const double* hit_bonus = a.sel.hit_bonus;
  if (q_f32) {
    // fp32 queries: warp w rounds query w of the group (quantize_rows_kernel's arithmetic: per-lane sum of squares over
    // d = lane, lane + 32, ..., shuffle tree, v * (1 / sqrt(ss)), round to nearest even) into shared memory; the bf16
    // rows also go to global memory once per group for the selection stage
    if (warp < nq) {
      const float* src = q_f32 + (size_t)(q0 + warp) * D;
      float scale = 1.f;
      if (a.q_normalize) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { const float t = src[d]; ss = fmaf(t, t, ss); }
        ss = warp_sum(ss);
        scale = 1.f / sqrtf(ss);
      }
      for (int d = lane; d < D; d += 32) {
        const uint16_t b16 = f32_to_bf16_rne(a.q_normalize ? src[d] * scale : src[d]);
        qs[(size_t)warp * D + d] = b16;
        if (blockIdx.x == 0) const_cast<uint16_t*>(a.s.q)[(size_t)(q0 + warp) * D + d] = b16;
      }
    }
    __syncthreads();
  }
  const uint16_t* qsrc = q_f32 ? qs : a.s.q + (size_t)q0 * D;
  float qr[QB][CH][8];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int chunk = lane + 32 * c;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (qq < nq && chunk < nchunk) v = *reinterpret_cast<const uint4*>(qsrc + (size_t)qq * D + chunk * 8);
      qr[qq][c][0] = bf16_lo(v.x); qr[qq][c][1] = bf16_hi(v.x);
      qr[qq][c][2] = bf16_lo(v.y); qr[qq][c][3] = bf16_hi(v.y);
      qr[qq][c][4] = bf16_lo(v.z); qr[qq][c][5] = bf16_hi(v.z);
      qr[qq][c][6] = bf16_lo(v.w); qr[qq][c][7] = bf16_hi(v.w);
    }
  }
  float wg[QB][2];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
    wg[qq][0] = a.s.w[0]; wg[qq][1] = a.s.w[1];
    if (a.s.wq[0] && qq < nq) { wg[qq][0] = a.s.wq[0][q0 + qq]; wg[qq][1] = a.s.wq[1][q0 + qq]; }
  }
  __syncthreads();

  constexpr int NV = kStreamRW * QB;                          // (row, query) sums per warp and step
  const int my_val = stream_value_of<NV>(lane);               // the sum this lane holds after the reduction
  const bool my_turn = (lane & (32 / NV - 1)) == 0;           // one lane per sum appends
  for (int it = 0; it < niter; ++it) {
    const int64_t row_w = r0 + ((int64_t)it * kStreamWarps + warp) * kStreamRW;       // first row of this warp
    // every load of the step goes out before the first use
    uint4 x[kStreamRW][2][CH];
#pragma unroll
    for (int r = 0; r < kStreamRW; ++r)
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int chunk = lane + 32 * c;
          x[r][g][c] = (g < a.s.G && chunk < nchunk && row_w + r < r1)
                           ? ldg_stream(a.s.gal[g] + (size_t)(row_w + r) * D + chunk * 8) : make_uint4(0, 0, 0, 0);
        }
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (g < a.s.G) {
        float acc[kStreamRW][QB][2];                            // two independent chains per (row, query)
#pragma unroll
        for (int r = 0; r < kStreamRW; ++r)
#pragma unroll
          for (int qq = 0; qq < QB; ++qq) { acc[r][qq][0] = 0.f; acc[r][qq][1] = 0.f; }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
          for (int r = 0; r < kStreamRW; ++r) {
            const uint4 y = x[r][g][c];
            const float x0 = bf16_lo(y.x), x1 = bf16_hi(y.x), x2 = bf16_lo(y.y), x3 = bf16_hi(y.y);
            const float x4 = bf16_lo(y.z), x5 = bf16_hi(y.z), x6 = bf16_lo(y.w), x7 = bf16_hi(y.w);
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
              float t0 = acc[r][qq][0], t1 = acc[r][qq][1];
              t0 = fmaf(x0, qr[qq][c][0], t0); t1 = fmaf(x1, qr[qq][c][1], t1);
              t0 = fmaf(x2, qr[qq][c][2], t0); t1 = fmaf(x3, qr[qq][c][3], t1);
              t0 = fmaf(x4, qr[qq][c][4], t0); t1 = fmaf(x5, qr[qq][c][5], t1);
              t0 = fmaf(x6, qr[qq][c][6], t0); t1 = fmaf(x7, qr[qq][c][7], t1);
              acc[r][qq][0] = t0; acc[r][qq][1] = t1;
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kStreamRW; ++r)
#pragma unroll
          for (int qq = 0; qq < QB; ++qq)
            v[r * QB + qq] = fmaf(wg[qq][g], acc[r][qq][0] + acc[r][qq][1], v[r * QB + qq]);
      }
    }
    const float sc = stream_reduce<NV>(v, lane);
    {
      const int r = my_val / QB, qq = my_val % QB;
      const int64_t row = row_w + r;
      if (my_turn && row < r1 && qq < nq) {
        const uint64_t key = make_key(sc, (uint32_t)row);
        if (key > s_thr[qq]) {
          const unsigned int slot = atomicAdd(&s_cnt[qq], 1u);
          cand[(size_t)qq * kStreamCand + slot] = key;              // the fill check below keeps slot < kStreamCand
        }
      }
    }
    if ((it + 1) % kStreamCheck == 0 && it + 1 < niter) {
      // at most kStreamCheck * kStreamRows keys per query arrived since the last look
      __syncthreads();
      bool cut = false;
#pragma unroll
      for (int qq = 0; qq < QB; ++qq) cut = cut || s_cnt[qq] > (unsigned)(kStreamCand - kStreamCheck * kStreamRows - kStreamRows);
      if (cut) {
        for (int qq = 0; qq < nq; ++qq) stream_cut(cand + (size_t)qq * kStreamCand, best + (size_t)qq * kMaxKSel, (int)s_cnt[qq], K);
        __syncthreads();
        for (int qq = 0; qq < nq; ++qq) {
          const int n = (int)s_cnt[qq], m = n < K ? n : K;
          for (int i = threadIdx.x; i < m; i += kStreamThreads) cand[(size_t)qq * kStreamCand + i] = best[(size_t)qq * kMaxKSel + i];
        }
        __syncthreads();
        if (threadIdx.x < nq) {
          const int n = (int)s_cnt[threadIdx.x];
          if (n >= K) { s_thr[threadIdx.x] = best[(size_t)threadIdx.x * kMaxKSel + K - 1]; s_cnt[threadIdx.x] = (unsigned)K; }
        }
      }
      __syncthreads();
    }
  }
  // final cut: this CTA's K best per query, in order, to its list
  __syncthreads();
  for (int qq = 0; qq < nq; ++qq) {
    const int n = (int)s_cnt[qq];
    uint64_t* dst = a.s.part_keys + ((size_t)blockIdx.x * a.s.Q + (q0 + qq)) * K;
    const uint64_t* cq = cand + (size_t)qq * kStreamCand;
    for (int i = threadIdx.x; i < n; i += kStreamThreads) {
      const uint64_t xk = cq[i];
      int r = 0;
      for (int j = 0; j < n; ++j) r += cq[j] > xk ? 1 : 0;
      if (r < K) dst[r] = xk;
    }
    for (int i = n + threadIdx.x; i < K; i += kStreamThreads) dst[i] = 0;     // fewer than K rows seen: empty tail
  }

  // KG hits are known before the scan: their canonical final scores are computed here, one hit per warp, spread over
  // the CTAs (hit j of the group -> CTA j mod P), so the selection of the last CTA finds them ready instead of
  // re-scoring ~20 extra rows on its own (batch 1, 43 k rows: 6 us of a 16 us selection)
  if (a.sel.hit_score && hit_rowptr) {
    const long long hb = hit_rowptr[q0], he = hit_rowptr[q0 + nq];
    for (long long j = hb + blockIdx.x + (long long)warp * P; j < he; j += (long long)P * kStreamWarps) {
      int qq = 0;
      while (qq + 1 < nq && hit_rowptr[q0 + qq + 1] <= j) ++qq;
      const int32_t col = hit_col[j];
      double f = __longlong_as_double(0x7ff8000000000000ll);
      if (col >= 0 && (int64_t)col < a.s.M) {                  // warp-uniform
        CanonQueryT<NP> cq;
        cq.load(qsrc + (size_t)qq * D, D, lane);
        CanonRow<NP> ra, rb;
        const size_t off = (size_t)col * D;
        ra.load(a.s.gal[0] + off, D, lane);
        rb.load(a.s.gal[a.s.G > 1 ? 1 : 0] + off, D, lane);
        double sa, sb;
        cq.dot2_lane0(ra, rb, D, lane, sa, sb);
        const double wa = a.sel.wq[0] ? a.sel.wq[0][q0 + qq] : a.sel.w[0], wb = a.sel.wq[0] ? a.sel.wq[1][q0 + qq] : a.sel.w[1];
        f = canon_fuse(sa, sb, a.s.G > 1, wa, wb, a.sel.alpha, hit_bonus[j], true);
      }
      if (lane == 0) const_cast<double*>(a.sel.hit_score)[j - hit_rowptr[0]] = f;     // slots count from the CSR's first entry
    }
  }

  // ------------------------------------------------------------------- last CTA of the group runs the selection
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&a.done[group], 1u);
    s_last = prev == (unsigned)(P - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (a.stamps && group == 0 && threadIdx.x == 0) a.stamps[1] = (long long)ptx::globaltimer();
  for (int qq = 0; qq < nq; ++qq) select_query<NP, kStreamWarps>(a.sel, q0 + qq, smem);
  if (threadIdx.x == 0) {
    a.done[group] = 0;                                            // ready for the next launch
    if (a.stamps && group == 0) a.stamps[2] = (long long)ptx::globaltimer();
  }
}

}  // namespace kemr
