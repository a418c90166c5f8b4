// Candidate selection, canonical binary64 re-scoring and rank finalisation.
//
// The scan kernels (scan_warp.cuh, scan_mma.cuh) only SELECT: per query they leave P sorted
// lists of K fp32-scored candidates.  The kernels here decide every returned index, score and
// rank on canonical binary64 scores (common.cuh::canon_dot_warp), merge the sparse KG-hit side
// path, and emit the per-query certificate.
#pragma once
#include "common.cuh"

namespace kemr {

struct SelectArgs {
  const uint64_t* part_keys;   // [P][Q][Kp]
  int P, Q, K;                 // K = candidates re-scored per query
  int Kp;                      // entries per part list (Kp <= K)
  const uint16_t* q;
  const uint16_t* gal[2];
  int G, D;
  int64_t M;
  double w[2];
  double alpha;
  const int64_t* hit_rowptr;   // [Q+1] or null
  const int32_t* hit_col;
  const double* hit_bonus;
  int k;
  double eps;
  int64_t idx_base;
  double* out_score64;         // [Q][k]
  float* out_score32;          // [Q][k] or null
  int64_t* out_idx;            // [Q][k]
  int32_t* out_flags;          // [Q]
  int max_cand;                // K + max hits per query
};

constexpr int kSelectThreads = 256;
constexpr int kSelectWarps = kSelectThreads / 32;
constexpr int kCountMax = 4096;          // up to this many candidate keys are ranked by counting

// dynamic smem: keys[max(kSelectWarps*K, min(P*K, kCountMax))] u64 | cand_score[max_cand] f64 |
//   cand_bonus[max_cand] f64 | cand_row[max_cand] i32 | cand_has[max_cand] u8 (padded) |
//   qrow[D] bf16 | rows[kSelectWarps][G][D] bf16
inline size_t select_key_slots(int P, int Kp, int K) {
  const size_t all = (size_t)P * Kp, per_warp = (size_t)kSelectWarps * K;
  return all <= (size_t)kCountMax ? (all > per_warp ? all : per_warp) : per_warp;
}
inline size_t select_smem_bytes(int P, int Kp, int K, int max_cand, int G, int D) {
  size_t b = select_key_slots(P, Kp, K) * 8 + (size_t)max_cand * (8 + 8 + 4);
  b += ((size_t)max_cand + 15) & ~(size_t)15;
  b += (size_t)D * 2 + (size_t)kSelectWarps * G * D * 2;
  return (b + 15) & ~(size_t)15;
}

// canonical dot from rows staged in shared memory (same order as common.cuh::canon_dot_warp)
__device__ __forceinline__ double canon_dot_smem(const uint16_t* a, const uint16_t* b, int D, int lane) {
  double acc = 0.0;
  for (int d = lane; d < D; d += 32)
    acc = fma((double)bf16_to_f32(a[d]), (double)bf16_to_f32(b[d]), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, o));
  return __shfl_sync(0xffffffffu, acc, 0);
}

__global__ void __launch_bounds__(kSelectThreads) select_rescore_kernel(SelectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K, Kp = a.Kp;
  const int n_all = a.P * Kp;
  const bool by_count = n_all <= kCountMax;
  // carve-up (16-byte aligned pieces first): qrow | rows | keys | cand_score | cand_bonus | cand_row | cand_has
  uint16_t* s_q = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* s_rows = s_q + a.D;
  uint64_t* lists = reinterpret_cast<uint64_t*>(s_rows + (size_t)kSelectWarps * a.G * a.D);
  const size_t per_warp = (size_t)kSelectWarps * K;
  const size_t key_slots = by_count ? ((size_t)n_all > per_warp ? (size_t)n_all : per_warp) : per_warp;
  double* cand_score = reinterpret_cast<double*>(lists + key_slots);
  double* cand_bonus = cand_score + a.max_cand;
  int32_t* cand_row = reinterpret_cast<int32_t*>(cand_bonus + a.max_cand);
  unsigned char* cand_has = reinterpret_cast<unsigned char*>(cand_row + a.max_cand);
  __shared__ int s_nsel, s_extra;
  __shared__ uint64_t s_sel[kMaxKSel];
  __shared__ unsigned long long s_bound;     // largest key any stage rejected (0 = nothing rejected)

  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // stage the query row (coalesced 16-byte loads)
  for (int c = threadIdx.x; c < (a.D >> 3); c += blockDim.x)
    reinterpret_cast<uint4*>(s_q)[c] = reinterpret_cast<const uint4*>(a.q + (size_t)qi * a.D)[c];

  // ---- A. the K best fp32 candidates over all parts
  if (by_count) {
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_sel[i] = 0;
    if (threadIdx.x == 0) s_bound = 0;
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
      const int p = i / Kp, j = i - p * Kp;
      lists[i] = a.part_keys[((size_t)p * a.Q + qi) * Kp + j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
      const uint64_t x = lists[i];
      if (!x) continue;
      int r = 0;
      for (int j = 0; j < n_all; ++j) r += lists[j] > x ? 1 : 0;
      if (r < K) s_sel[r] = x;                            // keys are distinct -> ranks are distinct
      // rejected here (rank >= K), or last entry of a full part list (the part rejected rows below it)
      if (r >= K || (i % Kp) == Kp - 1) atomicMax(&s_bound, (unsigned long long)x);
    }
    __syncthreads();
    if (warp == 0) {
      int n = 0;
      for (int i = lane; i < K; i += 32) n += (s_sel[i] != 0);
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (lane == 0) { s_nsel = n; s_extra = 0; }
    }
  } else {
    uint64_t* mine = lists + (size_t)warp * K;
    for (int i = lane; i < K; i += 32) mine[i] = 0;
    __syncwarp();
    uint64_t thr = 0;
    if (threadIdx.x == 0) s_bound = 0;
    __syncthreads();
    unsigned long long bnd = 0;
    for (int p = warp; p < a.P; p += kSelectWarps) {
      const uint64_t* src = a.part_keys + ((size_t)p * a.Q + qi) * Kp;
      const uint64_t last = src[Kp - 1];
      if (last > bnd) bnd = last;                 // a full part list rejected rows below its last key
      for (int i = 0; i < Kp; ++i) {
        const uint64_t x = src[i];
        if (x <= thr) { if (x > bnd) bnd = x; break; }   // sorted: this and the rest are rejected here
        warp_list_insert(mine, K, x, lane);
        thr = mine[K - 1];
      }
    }
    if (lane == 0 && bnd) atomicMax(&s_bound, bnd);
    __syncthreads();
    if (warp == 0) {
      thr = mine[K - 1];
      for (int w2 = 1; w2 < kSelectWarps; ++w2) {
        const uint64_t* other = lists + (size_t)w2 * K;
        for (int i = 0; i < K; ++i) {
          const uint64_t x = other[i];
          if (x <= thr) break;
          warp_list_insert(mine, K, x, lane);
          thr = mine[K - 1];
        }
      }
      // keys displaced from / never admitted to the final list are bounded by its last key
      if (lane == 0 && mine[K - 1]) atomicMax(&s_bound, (unsigned long long)mine[K - 1]);
      int n = 0;
      for (int i = lane; i < K; i += 32) { s_sel[i] = mine[i]; n += (mine[i] != 0); }
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (lane == 0) { s_nsel = n; s_extra = 0; }
    }
  }
  __syncthreads();
  const int nsel = s_nsel;
  const uint64_t* sel = s_sel;

  // ---- B. candidate table = scan candidates U KG hits
  for (int i = threadIdx.x; i < nsel; i += blockDim.x) {
    cand_row[i] = (int32_t)key_row(sel[i]);
    cand_bonus[i] = 0.0;
    cand_has[i] = 0;
  }
  __syncthreads();
  if (a.hit_rowptr) {
    const int64_t h0 = a.hit_rowptr[qi], h1 = a.hit_rowptr[qi + 1];
    for (int64_t h = h0 + threadIdx.x; h < h1; h += blockDim.x) {
      const int32_t col = a.hit_col[h];
      if (col < 0 || (int64_t)col >= a.M) continue;
      int found = -1;
      for (int i = 0; i < nsel; ++i) if (cand_row[i] == col) { found = i; break; }
      if (found < 0) {
        found = nsel + atomicAdd(&s_extra, 1);
        if (found >= a.max_cand) continue;      // cannot happen when max_hits_per_query is honest
        cand_row[found] = col;
      }
      cand_bonus[found] = a.hit_bonus[h];
      cand_has[found] = 1;
    }
  }
  __syncthreads();
  const int n = min(nsel + s_extra, a.max_cand);

  // ---- C. canonical re-scoring, one warp per candidate; rows are fetched with coalesced
  //         16-byte loads into shared memory, then accumulated in the canonical lane order
  uint16_t* my_rows = s_rows + (size_t)warp * a.G * a.D;
  const int nchunk = a.D >> 3;
  for (int c = warp; c < n; c += kSelectWarps) {
    const size_t off = (size_t)cand_row[c] * a.D;
    for (int g = 0; g < a.G; ++g)
      for (int ch = lane; ch < nchunk; ch += 32)
        reinterpret_cast<uint4*>(my_rows + (size_t)g * a.D)[ch] = ldg_stream(a.gal[g] + off + ch * 8);
    __syncwarp();
    const double sa = canon_dot_smem(s_q, my_rows, a.D, lane);
    const double sb = a.G > 1 ? canon_dot_smem(s_q, my_rows + a.D, a.D, lane) : 0.0;
    if (lane == 0)
      cand_score[c] = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, cand_bonus[c], cand_has[c] != 0);
    __syncwarp();
  }
  __syncthreads();

  // ---- D. order by (score desc, row asc) by counting; write the first k
  __shared__ double s_kth;
  if (threadIdx.x == 0) s_kth = -INFINITY;
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double sc = cand_score[c];
    const int32_t rc = cand_row[c];
    int r = 0;
    for (int j = 0; j < n; ++j) r += ahead64(cand_score[j], cand_row[j], sc, rc) ? 1 : 0;
    if (r < a.k) {
      const size_t o = (size_t)qi * a.k + r;
      a.out_score64[o] = sc;
      if (a.out_score32) a.out_score32[o] = (float)sc;
      a.out_idx[o] = a.idx_base + rc;
      if (r == a.k - 1) s_kth = sc;
    }
  }
  for (int r = n + threadIdx.x; r < a.k; r += blockDim.x) {
    const size_t o = (size_t)qi * a.k + r;
    a.out_score64[o] = -INFINITY;
    if (a.out_score32) a.out_score32[o] = -INFINITY;
    a.out_idx[o] = -1;
  }
  __syncthreads();

  // ---- E. certificate: nothing the scan rejected can reach the k-th canonical score
  if (threadIdx.x == 0) {
    int flag = 0;
    if (s_bound) {                              // something was rejected on its fp32 score
      const double bound = (double)key_score((uint64_t)s_bound) + a.eps * (1.0 + 1.0 / 64.0);
      const double reach = a.alpha * bound + 1e-300;
      if (!(n >= a.k && s_kth > reach)) flag = 1;
    }
    a.out_flags[qi] = flag;
  }
}

// ------------------------------------------------------------------ canonical pair scores
__global__ void score_pairs_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ ga,
                                   const uint16_t* __restrict__ gb, int D, double wa, double wb,
                                   double alpha, const int32_t* __restrict__ pq,
                                   const int64_t* __restrict__ prow, const double* __restrict__ pbonus,
                                   int64_t n, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = wid; i < n; i += nw) {
    const uint16_t* qrow = q + (size_t)pq[i] * D;
    const size_t off = (size_t)prow[i] * D;
    const double sa = canon_dot_warp(qrow, ga + off, D, lane);
    const double sb = gb ? canon_dot_warp(qrow, gb + off, D, lane) : 0.0;
    if (lane == 0)
      out[i] = canon_fuse(sa, sb, gb != nullptr, wa, wb, alpha, pbonus ? pbonus[i] : 0.0, pbonus != nullptr);
  }
}

// ------------------------------------------------------------------ rank path
// band around the target's clip-level score inside which the fp32 scan cannot decide
__global__ void rank_band_kernel(const double* __restrict__ t, double alpha, double eps, int Q,
                                 float* __restrict__ lo, float* __restrict__ hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  const double c = t[i] / alpha;
  const double e = eps * (1.0 + 1.0 / 64.0) + fabs(c) * 1e-12;
  lo[i] = __double2float_rd(c - e);
  hi[i] = __double2float_ru(c + e);
}

__global__ void rank_sum_parts_kernel(const int32_t* __restrict__ part_count, int P, int Q,
                                      const unsigned int* __restrict__ amb_counter, unsigned int amb_cap,
                                      unsigned long long* __restrict__ count, int32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  unsigned long long s = 0;
  for (int p = 0; p < P; ++p) s += (unsigned long long)part_count[(size_t)p * Q + i];
  count[i] = s;
  flags[i] = (*amb_counter > amb_cap) ? 2 : 0;
}

struct RankFixArgs {
  const uint16_t* q;
  const uint16_t* gal[2];
  int G, D;
  double w[2];
  double alpha;
  const double* t;            // [Q]
  const int64_t* t_gidx;      // [Q]
  int64_t idx_base;
  unsigned long long* count;  // [Q]
};

// ambiguous rows: decide on canonical scores (KG bonus deliberately ignored here; the hit
// correction below accounts for it)
__global__ void rank_amb_kernel(RankFixArgs a, const uint32_t* __restrict__ amb_q,
                                const uint32_t* __restrict__ amb_row,
                                const unsigned int* __restrict__ amb_counter, unsigned int amb_cap) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned int n = min(*amb_counter, amb_cap);
  for (int64_t i = wid; i < n; i += nw) {
    const int qi = (int)amb_q[i];
    const int64_t row = amb_row[i];
    const uint16_t* qrow = a.q + (size_t)qi * a.D;
    const size_t off = (size_t)row * a.D;
    const double sa = canon_dot_warp(qrow, a.gal[0] + off, a.D, lane);
    const double sb = a.G > 1 ? canon_dot_warp(qrow, a.gal[1] + off, a.D, lane) : 0.0;
    if (lane == 0) {
      const double f = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, 0.0, false);
      if (ahead64(f, a.idx_base + row, a.t[qi], a.t_gidx[qi])) atomicAdd(&a.count[qi], 1ull);
    }
  }
}

// KG hits: replace each hit row's un-boosted verdict by its boosted one
__global__ void rank_hits_kernel(RankFixArgs a, int Q, int64_t M, const int64_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ col, const double* __restrict__ bonus) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    const uint16_t* qrow = a.q + (size_t)qi * a.D;
    const int64_t h0 = rowptr[qi], h1 = rowptr[qi + 1];
    for (int64_t h = h0 + warp; h < h1; h += nwarp) {
      const int64_t row = col[h];
      if (row < 0 || row >= M) continue;
      const size_t off = (size_t)row * a.D;
      const double sa = canon_dot_warp(qrow, a.gal[0] + off, a.D, lane);
      const double sb = a.G > 1 ? canon_dot_warp(qrow, a.gal[1] + off, a.D, lane) : 0.0;
      if (lane == 0) {
        const double f0 = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, 0.0, false);
        const double f1 = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, bonus[h], true);
        const int d = (int)ahead64(f1, a.idx_base + row, a.t[qi], a.t_gidx[qi]) -
                      (int)ahead64(f0, a.idx_base + row, a.t[qi], a.t_gidx[qi]);
        if (d) atomicAdd(&a.count[qi], (unsigned long long)(long long)d);
      }
    }
  }
}

// ------------------------------------------------------------------ post-all-gather merge
__global__ void merge_topk_kernel(const double* __restrict__ in_s, const int64_t* __restrict__ in_i,
                                  int R, int Q, int k, double* __restrict__ out_s,
                                  int64_t* __restrict__ out_i) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s = reinterpret_cast<double*>(smem_raw);
  int64_t* ix = reinterpret_cast<int64_t*>(s + (size_t)R * k);
  const int qi = blockIdx.x;
  const int n = R * k;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const int r = c / k, j = c % k;
    const size_t o = ((size_t)r * Q + qi) * k + j;
    s[c] = in_s[o];
    ix[c] = in_i[o];
  }
  __syncthreads();
  __shared__ int s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  int myvalid = 0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    if (ix[c] < 0) continue;
    ++myvalid;
    int r = 0;
    for (int j = 0; j < n; ++j) r += (ix[j] >= 0 && ahead64(s[j], ix[j], s[c], ix[c])) ? 1 : 0;
    if (r < k) { out_s[(size_t)qi * k + r] = s[c]; out_i[(size_t)qi * k + r] = ix[c]; }
  }
  atomicAdd(&s_valid, myvalid);
  __syncthreads();
  for (int r = s_valid + threadIdx.x; r < k; r += blockDim.x) {
    out_s[(size_t)qi * k + r] = -INFINITY;
    out_i[(size_t)qi * k + r] = -1;
  }
}

}  // namespace kemr
