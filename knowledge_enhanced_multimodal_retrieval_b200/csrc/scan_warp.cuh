// Warp-dot scan for small query batches in the COUNT and DENSE modes (the top-k search of small batches is the fused
// streaming kernel of scan_stream.cuh).
//
// One warp owns one gallery row at a time: the row's D bf16 values are fetched with fully
// coalesced 16-byte loads (lane l takes chunk l, l+32, ...), multiplied against query values that
// live in registers as fp32, and reduced with warp shuffles.  The T2I/T2T weights are applied to
// the per-lane partial sums before the (single) reduction, so two galleries cost one shuffle
// tree.  The epilogue is one of
//   COUNT : number of rows scoring above a per-query threshold, rows inside the +-eps band go
//           to an "ambiguous" list that is re-scored in binary64 later (ranks for Recall@K / MRR);
//   DENSE : fp32 scores written out (compatibility / diagnostics).
// Algorithmic traffic is G*M*D*2 bytes per query group; HBM-bound (roofline: MEASURED_PEAKS hbm_gbs).
#pragma once
#include "common.cuh"

namespace kemr {

enum ScanMode { kModeTopk = 0, kModeCount = 1, kModeDense = 2 };

struct ScanArgs {
  const uint16_t* q;        // [Q][D]
  int Q;
  const uint16_t* gal[2];   // [M][D]
  int G;
  int64_t M;
  int D;
  float w[2];
  const float* wq[2];       // optional per-query fusion weights [Q] (gated fusion heads); null = the scalars w[]
  int mode;
  // TOPK
  int K;
  uint64_t* part_keys;      // [P][Q][K], descending, 0 = empty
  unsigned int* thr_pub;    // optional [Q rows], zero on entry: per query the largest K-th score any FULL list has published
                            // (order_f32 image, 0 = none) -- a pruning threshold shared by all lists of the query (scan_mma.cuh)
  // COUNT
  const float* band_lo;     // [Q]
  const float* band_hi;     // [Q]
  int32_t* part_count;      // [P][Q]
  uint32_t* amb_q;          // [amb_cap]
  uint32_t* amb_row;
  unsigned int* amb_counter;
  unsigned int amb_cap;
  // DENSE
  float* dense;
  int64_t ld;
};

constexpr int kWarpScanThreads = 256;
constexpr int kWarpScanWarps = kWarpScanThreads / 32;

template <int QB, int CH>
__global__ void __launch_bounds__(kWarpScanThreads, 2) scan_warp_kernel(ScanArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.y * QB;
  const int nq = min(QB, a.Q - q0);
  const int P = gridDim.x;
  const int64_t r0 = (a.M * (int64_t)blockIdx.x) / P;
  const int64_t r1 = (a.M * (int64_t)(blockIdx.x + 1)) / P;
  const int nchunk = a.D >> 3;

  // queries -> fp32 registers (lane l keeps the d-slices it will meet in every row)
  float qr[QB][CH][8];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int chunk = lane + 32 * c;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (qq < nq && chunk < nchunk)
        v = *reinterpret_cast<const uint4*>(a.q + (size_t)(q0 + qq) * a.D + chunk * 8);
      qr[qq][c][0] = bf16_lo(v.x); qr[qq][c][1] = bf16_hi(v.x);
      qr[qq][c][2] = bf16_lo(v.y); qr[qq][c][3] = bf16_hi(v.y);
      qr[qq][c][4] = bf16_lo(v.z); qr[qq][c][5] = bf16_hi(v.z);
      qr[qq][c][6] = bf16_lo(v.w); qr[qq][c][7] = bf16_hi(v.w);
    }
  }

  float wg[QB][2];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
    wg[qq][0] = a.w[0]; wg[qq][1] = a.w[1];
    if (a.wq[0] && qq < nq) { wg[qq][0] = a.wq[0][q0 + qq]; wg[qq][1] = a.wq[1][q0 + qq]; }
  }
  int32_t cnt[QB];
  float blo[QB], bhi[QB];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
    cnt[qq] = 0; blo[qq] = 0.f; bhi[qq] = 0.f;
    if (a.mode == kModeCount && qq < nq) { blo[qq] = a.band_lo[q0 + qq]; bhi[qq] = a.band_hi[q0 + qq]; }
  }
  // two rows in flight per warp
  for (int64_t row = r0 + warp; row < r1; row += 2 * kWarpScanWarps) {
    const int64_t rowB = row + kWarpScanWarps;
    const bool hasB = rowB < r1;
    uint4 va[2][2][CH];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (g < a.G) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int chunk = lane + 32 * c;
          const bool ok = chunk < nchunk;
          va[0][g][c] = ok ? ldg_stream(a.gal[g] + (size_t)row * a.D + chunk * 8) : make_uint4(0, 0, 0, 0);
          va[1][g][c] = (ok && hasB) ? ldg_stream(a.gal[g] + (size_t)rowB * a.D + chunk * 8)
                                     : make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr == 1 && !hasB) break;
      const int64_t rcur = rr ? rowB : row;
      float s[QB];
#pragma unroll
      for (int qq = 0; qq < QB; ++qq) s[qq] = 0.f;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g < a.G) {
          float acc[QB];
#pragma unroll
          for (int qq = 0; qq < QB; ++qq) acc[qq] = 0.f;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const uint4 v = va[rr][g][c];
            const float x0 = bf16_lo(v.x), x1 = bf16_hi(v.x), x2 = bf16_lo(v.y), x3 = bf16_hi(v.y);
            const float x4 = bf16_lo(v.z), x5 = bf16_hi(v.z), x6 = bf16_lo(v.w), x7 = bf16_hi(v.w);
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
              float t = acc[qq];
              t = fmaf(x0, qr[qq][c][0], t); t = fmaf(x1, qr[qq][c][1], t);
              t = fmaf(x2, qr[qq][c][2], t); t = fmaf(x3, qr[qq][c][3], t);
              t = fmaf(x4, qr[qq][c][4], t); t = fmaf(x5, qr[qq][c][5], t);
              t = fmaf(x6, qr[qq][c][6], t); t = fmaf(x7, qr[qq][c][7], t);
              acc[qq] = t;
            }
          }
#pragma unroll
          for (int qq = 0; qq < QB; ++qq) s[qq] = fmaf(wg[qq][g], acc[qq], s[qq]);
        }
      }
#pragma unroll
      for (int qq = 0; qq < QB; ++qq) {
        const float sc = warp_sum(s[qq]);
        if (qq >= nq) continue;
        if (a.mode == kModeCount) {
          if (sc > bhi[qq]) {
            cnt[qq]++;
          } else if (sc >= blo[qq] && lane == 0) {
            const unsigned int slot = atomicAdd(a.amb_counter, 1u);
            if (slot < a.amb_cap) { a.amb_q[slot] = (uint32_t)(q0 + qq); a.amb_row[slot] = (uint32_t)rcur; }
          }
        } else {
          if (lane == 0) a.dense[(size_t)(q0 + qq) * a.ld + rcur] = sc;
        }
      }
    }
  }

  if (a.mode == kModeCount) {
    __shared__ int32_t csum[QB];
    if (threadIdx.x < QB) csum[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int qq = 0; qq < QB; ++qq) if (cnt[qq]) atomicAdd(&csum[qq], cnt[qq]);
    }
    __syncthreads();
    if (threadIdx.x < nq) a.part_count[(size_t)blockIdx.x * a.Q + q0 + threadIdx.x] = csum[threadIdx.x];
  }
}

}  // namespace kemr
