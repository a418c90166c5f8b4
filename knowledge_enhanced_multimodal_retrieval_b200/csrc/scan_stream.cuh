// HBM-streaming warp-dot search for small, latency-bound query batches, scan + selection in ONE launch
// (north_star kernel 1b + 2 for batch 1..4; BASELINE config 3).
//
// One persistent CTA per SM owns a contiguous range of gallery rows.  A producer warp streams that range -- which is
// one contiguous byte range per gallery -- through a ring of shared-memory stages with bulk asynchronous copies
// (cp.async.bulk -> UBLKCP, completion on an mbarrier), so hundreds of KB per SM are in flight without costing a
// register.  Eight consumer warps take two rows of a 16-row stage each: conflict-free 16-byte shared-memory reads,
// fp32 FMAs in independent chains against the queries held in registers (fusion weights applied to the per-lane
// partial sums, so T2I + T2T cost one reduction), ONE transposed shuffle reduction for all (row, query) sums of the
// warp (2..8 values: 5..9 shuffles instead of 5 per value), and the lanes that end up holding a sum append (score,
// row) keys that beat the CTA's running threshold to a candidate buffer.
// Nothing is sorted while rows stream: the buffer is cut back to its K best by counting only when it fills (never,
// for a 43 k-row gallery) and once at the end, which leaves one K-entry list per CTA and query.
// The last CTA to finish (device-wide arrival counter) then runs the selection stage of select.cuh in the same
// launch: merge of the lists, KG hits, canonical binary64 re-scoring, final order, certificate.  No second launch,
// no score matrix, no host round trip.
// Algorithmic traffic: G*M*D*2 bytes per group of QB queries (roofline: MEASURED_PEAKS hbm_gbs).
#pragma once
#include "scan_mma.cuh"     // ptx:: mbarrier helpers
#include "select.cuh"

namespace kemr {

constexpr int kStreamConsumers = 8;                       // consumer warps
constexpr int kStreamRW = 2;                              // rows per consumer warp and stage
constexpr int kStreamRows = kStreamConsumers * kStreamRW; // rows per stage
constexpr int kStreamThreads = (kStreamConsumers + 1) * 32;
constexpr int kStreamWarps = kStreamConsumers + 1;
constexpr int kStreamCand = 1024;                         // candidate keys per query a CTA can hold between cuts
constexpr int kStreamCheck = 16;                          // stages between two looks at the buffer fill (<= 256 appends)
constexpr int kStreamMaxStages = 12;

struct StreamArgs {
  ScanArgs s;               // queries, galleries, weights; s.K = list length per CTA; s.part_keys [P][Q][K]
  SelectArgs sel;           // the selection stage (P = gridDim.x lists per query)
  unsigned int* done;       // [groups] arrival counters, zero on entry, reset by the last CTA
  long long* stamps;        // optional [4] globaltimer stamps of group 0: first CTA start, last scan end, select end
  int stages;               // ring depth
  unsigned int stage_bytes; // G * kStreamRows * D * 2
  unsigned int tail_off;    // offset of the barriers / candidate buffers behind max(ring, selection scratch)
};

inline size_t stream_tail_off(int stages, size_t stage_bytes, size_t select_bytes) {
  return (std::max((size_t)stages * stage_bytes, select_bytes) + 15) / 16 * 16;
}
inline size_t stream_smem_bytes(int stages, size_t stage_bytes, int QB, size_t select_bytes) {
  const size_t tail = (size_t)2 * kStreamMaxStages * 8 + (size_t)QB * kStreamCand * 8 + (size_t)QB * kMaxKSel * 8 + 256;
  return stream_tail_off(stages, stage_bytes, select_bytes) + tail + 128;
}

namespace ptx {
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers * 32) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
}  // namespace ptx

// keep the K best of cand[0..n) (distinct keys): ranks by counting, all consumer threads; returns through s_* arrays.
// Caller guarantees a consumer_bar() before (appends visible) and runs one after.
__device__ __forceinline__ void stream_cut(uint64_t* cand, uint64_t* best, int n, int K, int ctid) {
  for (int i = ctid; i < n; i += kStreamConsumers * 32) {
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
    constexpr int kFull = 0;
    (void)kFull;
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
__global__ void __launch_bounds__(kStreamThreads, 1) scan_stream_kernel(StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_stream[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_stream + 127) & ~(uintptr_t)127);
  unsigned char* tail = smem + a.tail_off;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kStreamMaxStages;
  uint64_t* cand = empty_bar + kStreamMaxStages;              // [QB][kStreamCand]
  uint64_t* best = cand + (size_t)QB * kStreamCand;           // [QB][kMaxKSel]
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
  const int nrows = (int)(r1 - r0);
  const int niter = (nrows + kStreamRows - 1) / kStreamRows;
  const int K = a.s.K;
  const uint32_t row_bytes = (uint32_t)a.s.D * 2u;
  const uint32_t gal_bytes = kStreamRows * row_bytes;          // one gallery's share of a stage

  if (a.stamps && group == 0 && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(a.stamps), ptx::globaltimer());
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], kStreamConsumers); }
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < QB) { s_cnt[threadIdx.x] = 0; s_thr[threadIdx.x] = 0; }
  __syncthreads();

  if (warp == kStreamConsumers) {
    // ================================================================= producer: bulk copies into the ring
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < niter; ++it) {
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      if (ptx::elect_one()) {
        const int64_t row = r0 + (int64_t)it * kStreamRows;
        const uint32_t n = (uint32_t)min((int64_t)kStreamRows, r1 - row);
        const uint32_t bytes = n * row_bytes;
        const uint32_t dst = ptx::smem_u32(smem + (size_t)stage * a.stage_bytes);
        ptx::mbar_expect_tx(&full_bar[stage], bytes * (uint32_t)a.s.G);
        ptx::bulk_load(dst, a.s.gal[0] + (size_t)row * a.s.D, bytes, &full_bar[stage]);
        if (a.s.G > 1) ptx::bulk_load(dst + gal_bytes, a.s.gal[1] + (size_t)row * a.s.D, bytes, &full_bar[stage]);
      }
      __syncwarp();
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ================================================================= consumers: one row of a stage per warp
    const int ctid = threadIdx.x;                               // 0..255
    const int nchunk = a.s.D >> 3;
    float qr[QB][CH][8];
#pragma unroll
    for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int chunk = lane + 32 * c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (qq < nq && chunk < nchunk) v = *reinterpret_cast<const uint4*>(a.s.q + (size_t)(q0 + qq) * a.s.D + chunk * 8);
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
    constexpr int NV = kStreamRW * QB;                          // (row, query) sums per warp and stage
    const int my_val = stream_value_of<NV>(lane);               // the sum this lane holds after the reduction
    const bool my_turn = (lane & (32 / NV - 1)) == 0;           // one lane per sum appends
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < niter; ++it) {
      ptx::mbar_wait(&full_bar[stage], phase);
      const int64_t row_w = r0 + (int64_t)it * kStreamRows + (int64_t)warp * kStreamRW;     // first row of this warp
      float v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = 0.f;
      if (row_w < r1) {
        const unsigned char* sw = smem + (size_t)stage * a.stage_bytes + (size_t)warp * kStreamRW * row_bytes;
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
              const int chunk = lane + 32 * c;
              if (chunk < nchunk) {
#pragma unroll
                for (int r = 0; r < kStreamRW; ++r) {
                  const uint4 x = *reinterpret_cast<const uint4*>(sw + (size_t)g * gal_bytes + (size_t)r * row_bytes + (size_t)chunk * 16);
                  const float x0 = bf16_lo(x.x), x1 = bf16_hi(x.x), x2 = bf16_lo(x.y), x3 = bf16_hi(x.y);
                  const float x4 = bf16_lo(x.z), x5 = bf16_hi(x.z), x6 = bf16_lo(x.w), x7 = bf16_hi(x.w);
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
            }
#pragma unroll
            for (int r = 0; r < kStreamRW; ++r)
#pragma unroll
              for (int qq = 0; qq < QB; ++qq)
                v[r * QB + qq] = fmaf(wg[qq][g], acc[r][qq][0] + acc[r][qq][1], v[r * QB + qq]);
          }
        }
      }
      // this warp is done reading the stage: hand it back before the reduction
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
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
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
      if ((it + 1) % kStreamCheck == 0 && it + 1 < niter) {
        // at most kStreamCheck * kStreamRows keys per query arrived since the last look
        ptx::consumer_bar();
        bool cut = false;
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) cut = cut || s_cnt[qq] > (unsigned)(kStreamCand - kStreamCheck * kStreamRows - kStreamRows);
        if (cut) {
          for (int qq = 0; qq < nq; ++qq)
            stream_cut(cand + (size_t)qq * kStreamCand, best + (size_t)qq * kMaxKSel, (int)s_cnt[qq], K, ctid);
          ptx::consumer_bar();
          for (int qq = 0; qq < nq; ++qq) {
            const int n = (int)s_cnt[qq], m = n < K ? n : K;
            for (int i = ctid; i < m; i += kStreamConsumers * 32) cand[(size_t)qq * kStreamCand + i] = best[(size_t)qq * kMaxKSel + i];
          }
          ptx::consumer_bar();
          if (ctid < nq) {
            const int n = (int)s_cnt[ctid];
            if (n >= K) { s_thr[ctid] = best[(size_t)ctid * kMaxKSel + K - 1]; s_cnt[ctid] = (unsigned)K; }
          }
          ptx::consumer_bar();
        }
      }
    }
    // final cut: this CTA's K best per query, in order, to its list
    ptx::consumer_bar();
    for (int qq = 0; qq < nq; ++qq) {
      const int n = (int)s_cnt[qq];
      uint64_t* dst = a.s.part_keys + ((size_t)blockIdx.x * a.s.Q + (q0 + qq)) * K;
      const uint64_t* cq = cand + (size_t)qq * kStreamCand;
      for (int i = ctid; i < n; i += kStreamConsumers * 32) {
        const uint64_t x = cq[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += cq[j] > x ? 1 : 0;
        if (r < K) dst[r] = x;
      }
      for (int i = n + ctid; i < K; i += kStreamConsumers * 32) dst[i] = 0;     // fewer than K rows seen: empty tail
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
