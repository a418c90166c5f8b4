// Candidate selection, canonical binary64 re-scoring and rank finalisation.
//
// The scan kernels (scan_warp.cuh, scan_mma.cuh) only SELECT: per query they leave P sorted
// lists of K fp32-scored candidates.  The kernels here decide every returned index, score and
// rank on canonical binary64 scores (common.cuh::canon_dot_warp), merge the sparse KG-hit side
// path, and emit the per-query certificate.
#pragma once
#include "common.cuh"

namespace kemr {

constexpr int kMaxPeers = 8;             // ranks of one NVSwitch box

struct SelectArgs {
  const uint64_t* part_keys;   // [P][Q][Kp]
  int P, Q, K;                 // K = candidates re-scored per query
  int Kp;                      // entries per part list (Kp <= K)
  const uint16_t* q;
  const uint16_t* gal[2];
  int G, D;
  int64_t M;
  double w[2];
  const double* wq[2];         // optional per-query fusion weights [Q]; null = the scalars w[]
  double alpha;
  const int64_t* hit_rowptr;   // [Q+1] or null
  const int32_t* hit_col;
  const double* hit_bonus;
  const double* hit_score;     // optional [nnz], entry h - hit_rowptr[0]: canonical FINAL score (bonus included) of CSR entry h, computed ahead of
                               // the selection (fused streaming search: by the CTAs that finish their scan early)
  int k;
  double eps;
  int64_t idx_base;
  double* out_score64;         // [Q][k]
  float* out_score32;          // [Q][k] or null
  int64_t* out_idx;            // [Q][k]
  int32_t* out_flags;          // [Q]
  int max_cand;                // K + max hits per query
  int key_slots;               // select_key_slots(Kp, K)
  // result exchange over NVLink peer memory (kemr_peer_*): when n_peer > 0 the k result rows of a query are ALSO stored
  // into every rank's gather buffer (slot of this rank) and a per-query flag is released there -- the all-gather of
  // the row-sharded search is fused into this kernel (no collective launch in the step)
  int n_peer;
  unsigned char* peer_base[kMaxPeers];   // mapped base address of every rank's exchange buffer (own rank: local memory)
  long long peer_block;        // entries per (parity, rank) block = max_q * max_k
  long long peer_idx_region;   // byte offset of the idx region inside a buffer
  long long peer_flag_region;  // byte offset of the flag region
  int peer_max_q, peer_rank, peer_world;
  const unsigned int* peer_epoch;        // local device word: the step's epoch (bumped by kemr_peer_begin)
};

constexpr int kMaxParts = 4096;          // part lists per query the head pass can hold
constexpr int kSelSmallCand = 64;        // up to this many candidates a query is handled by a 4-warp CTA

// dynamic shared memory of select_kernel (8-byte pieces first):
//   heads[P] u64 | keys[pow2 >= K*Kp] u64 | selk[K] u64 | score[max_cand] f64 | bonus[max_cand] f64 |
//   lid[K] i32 | row[max_cand] i32 | has[max_cand] u8
inline size_t select_key_slots(int Kp, int K) {      // gathered keys, rounded up to a power of two (bitonic sort)
  size_t n = 32;
  while (n < (size_t)K * Kp) n <<= 1;
  return n;
}
constexpr int kSelCountMax = 256;        // more surviving keys than this are sorted instead of ranked by counting
inline size_t select_smem_bytes(int P, int Kp, int K, int max_cand) {
  size_t b = ((size_t)P + select_key_slots(Kp, K) + (size_t)K + 2 * (size_t)max_cand) * 8 + ((size_t)K + (size_t)max_cand) * 4;
  b += (size_t)max_cand;
  return (b + 15) & ~(size_t)15;
}

// One CTA of W warps per query.  Merges the query's P part lists into the K best fp32-scored candidates,
// adds the KG hits, re-scores every candidate canonically (binary64), orders them and emits the certificate.
//   A1  heads of the P lists: only the K lists with the largest heads can hold a top-K key;
//   A2  keys of those lists that are not below the K-th head (K keys are already >= it) are ranked by counting;
//   A3  candidates past the k-th whose fp32 score is more than 2 eps below the k-th fp32 score cannot reach the
//       top k (|fp32 - canonical| <= eps, a KG bonus only raises the k-th score): not re-scored, counted as
//       rejected for the certificate; KG hits among them come back through the hit list;
//   B   candidate table = scan candidates U KG hits;
//   C   warp w re-scores candidates w, w+W, ...; the rows of its next candidate are in flight meanwhile;
//   D   order by (score desc, row asc) by counting, write the first k;   E  certificate.
// Debug build: %globaltimer at the phase boundaries of query 0's selection (kemr_debug_select_stamps).
#ifdef KEMR_DEBUG
__device__ unsigned long long g_sel_stamps[16];
#define KEMR_SEL_STAMP(i)                                                                    \
  do { if (threadIdx.x == 0 && qi == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); g_sel_stamps[i] = t__; } } while (0)
#else
#define KEMR_SEL_STAMP(i) do { } while (0)
#endif

// Body of the selection for ONE query, run by a whole CTA of W warps (every thread of the CTA must call it; it uses
// __syncthreads).  `smem_raw`: select_smem_bytes(...) bytes of shared memory, 16-byte aligned.  Called by
// select_kernel (one CTA per query) and by the last CTA of the fused small-batch scan (scan_stream.cuh).
template <int NP, int W>
__device__ __forceinline__ void select_query(const SelectArgs& a, const int qi, unsigned char* smem_raw) {
  constexpr int T = W * 32;
  const int K = a.K, Kp = a.Kp, P = a.P, MC = a.max_cand;
  uint64_t* s_heads = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_keys = s_heads + P;
  uint64_t* s_selk = s_keys + a.key_slots;
  double* s_score = reinterpret_cast<double*>(s_selk + K);
  double* s_bonus = s_score + MC;
  int* s_lid = reinterpret_cast<int*>(s_bonus + MC);
  int32_t* s_row = s_lid + K;
  unsigned char* s_has = reinterpret_cast<unsigned char*>(s_row + MC);
  __shared__ unsigned long long s_bound;     // largest key any stage rejected (0 = nothing rejected)
  __shared__ unsigned long long s_cut;
  __shared__ int s_nlist, s_extra, s_nsurv, s_nsel, s_first;
  __shared__ long long s_h0, s_h1;           // this query's range of the KG-hit CSR, fetched while the lists merge
  __shared__ double s_kth;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();                                          // a previous query of this CTA is done with the static state
  KEMR_SEL_STAMP(0);

  CanonQueryT<NP> cq;
  cq.load(a.q + (size_t)qi * a.D, a.D, lane);

  // A1
  for (int p = tid; p < P; p += T) s_heads[p] = a.part_keys[((size_t)p * a.Q + qi) * Kp];
  for (int i = tid; i < K; i += T) { s_selk[i] = 0; s_lid[i] = -1; }
  if (tid == 0) { s_bound = 0; s_cut = 0; s_nlist = 0; s_extra = 0; s_nsurv = 0; s_nsel = 0; s_first = K; s_kth = -INFINITY; }
  if (tid == 32 % T && a.hit_rowptr) { s_h0 = a.hit_rowptr[qi]; s_h1 = a.hit_rowptr[qi + 1]; }
  __syncthreads();
  {
    // (a 64-bit atomicMax in shared memory is a compare-and-swap loop: hundreds of rejected heads hammering one word
    // cost 6 us of a batch-1 selection -- reduce per warp first, one atomic per warp)
    // Only the K largest heads matter.  A warp first finds the K-th largest head of the whole set by looking at its
    // own slice: every thread counts how many heads beat its head with four independent accumulators (the loop is
    // latency-bound otherwise).
    unsigned long long rej = 0;
    for (int p0 = 0; p0 < P; p0 += T) {
      const int p = p0 + tid;
      const uint64_t x = p < P ? s_heads[p] : 0;
      if (x) {
        int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        int j = 0;
        for (; j + 3 < P; j += 4) {
          r0 += s_heads[j] > x ? 1 : 0; r1 += s_heads[j + 1] > x ? 1 : 0;
          r2 += s_heads[j + 2] > x ? 1 : 0; r3 += s_heads[j + 3] > x ? 1 : 0;
        }
        for (; j < P; ++j) r0 += s_heads[j] > x ? 1 : 0;
        const int r = (r0 + r1) + (r2 + r3);
        if (r < K) { s_lid[r] = p; atomicAdd(&s_nlist, 1); }
        else if (x > rej) rej = x;                          // whole list rejected: nothing in it beats its head
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, rej, o);
      rej = other > rej ? other : rej;
    }
    if (lane == 0 && rej) atomicMax(&s_bound, rej);
  }
  __syncthreads();

  KEMR_SEL_STAMP(1);
  // A2
  const int nlist = s_nlist;                                // selected lists occupy s_lid[0..nlist), by head rank
  // A lower bound of the K-th best key: the first jc entries of every selected list are jc*nlist >= K keys, so the
  // K-th largest of them (jc = 1: the K-th head) is at most the K-th best overall; anything below it is out.
  uint64_t cut = 0;
  if (nlist >= K) {
    cut = s_heads[s_lid[K - 1]];
  } else if (nlist > 0) {
    const int jc = min(Kp, (K + nlist - 1) / nlist);
    const int nsub = nlist * jc;                            // <= K + nlist - 1 < 2K <= key slots
    for (int i = tid; i < nsub; i += T) {
      const int l = i / jc, j = i - l * jc;
      s_keys[i] = a.part_keys[((size_t)s_lid[l] * a.Q + qi) * Kp + j];
    }
    __syncthreads();
    for (int i = tid; i < nsub; i += T) {
      const uint64_t x = s_keys[i];
      if (!x) continue;
      int r = 0;
      for (int j = 0; j < nsub; ++j) r += s_keys[j] > x ? 1 : 0;
      if (r == K - 1) s_cut = x;                            // exists only if the subset holds K non-empty keys
    }
    __syncthreads();
    cut = s_cut;
  }
  {
    const int n2 = nlist * Kp;
    unsigned long long rej = 0;                             // largest key this thread saw rejected
    for (int i = tid; i < n2; i += T) {
      const int l = i / Kp, j = i - l * Kp;
      const uint64_t x = a.part_keys[((size_t)s_lid[l] * a.Q + qi) * Kp + j];
      if (!x) continue;
      if (x >= cut) s_keys[atomicAdd(&s_nsurv, 1)] = x;
      else if (x > rej) rej = x;
      if (j == Kp - 1 && x > rej) rej = x;                  // a full part list rejected rows below its last key
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, rej, o);
      rej = other > rej ? other : rej;
    }
    if (lane == 0 && rej) atomicMax(&s_bound, rej);
  }
  __syncthreads();
  const int ns = s_nsurv;
  if (ns <= kSelCountMax) {
    for (int i = tid; i < ns; i += T) {
      const uint64_t x = s_keys[i];
      int r = 0;
      for (int j = 0; j < ns; ++j) r += s_keys[j] > x ? 1 : 0;
      if (r < K) { s_selk[r] = x; atomicAdd(&s_nsel, 1); }  // keys are distinct -> ranks are distinct
      else atomicMax(&s_bound, (unsigned long long)x);
    }
  } else {
    // few lists with long tails (fewer lists than K: nothing could be pruned): bitonic sort, descending
    int n2p = 32;
    while (n2p < ns) n2p <<= 1;
    for (int i = ns + tid; i < n2p; i += T) s_keys[i] = 0;
    __syncthreads();
    for (int kk = 2; kk <= n2p; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2p; i += T) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const uint64_t x = s_keys[i], y = s_keys[ixj];
            const bool desc = (i & kk) == 0;
            if (desc ? (x < y) : (x > y)) { s_keys[i] = y; s_keys[ixj] = x; }
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < K && i < ns; i += T) s_selk[i] = s_keys[i];
    if (tid == 0) {
      s_nsel = ns < K ? ns : K;
      if (ns > K) atomicMax(&s_bound, (unsigned long long)s_keys[K]);     // largest key rejected here
    }
  }
  __syncthreads();

  KEMR_SEL_STAMP(2);
  // A3 (keys fill s_selk[0..nsel) in descending order, so everything from the first dropped key on drops)
  int nsel = s_nsel;
  if (nsel > a.k) {
    const double kth32 = (double)key_score(s_selk[a.k - 1]) - 2.0 * a.eps * (1.0 + 1.0 / 64.0);
    for (int i = a.k + tid; i < nsel; i += T)
      if ((double)key_score(s_selk[i]) < kth32) atomicMin(&s_first, i);
    __syncthreads();
    if (s_first < nsel) {
      nsel = s_first;
      if (tid == 0) atomicMax(&s_bound, (unsigned long long)s_selk[nsel]);
    }
  }

  KEMR_SEL_STAMP(3);
  // B
  for (int i = tid; i < nsel; i += T) { s_row[i] = (int32_t)key_row(s_selk[i]); s_bonus[i] = 0.0; s_has[i] = 0; }
  __syncthreads();
  if (a.hit_rowptr) {
    const int64_t h0 = s_h0, h1 = s_h1;
    for (int64_t h = h0 + tid; h < h1; h += T) {
      const int32_t col = a.hit_col[h];
      if (col < 0 || (int64_t)col >= a.M) continue;
      int found = -1;
      for (int i = 0; i < nsel; ++i) if (s_row[i] == col) { found = i; break; }
      if (found < 0) {
        found = nsel + atomicAdd(&s_extra, 1);
        if (found >= MC) continue;                          // cannot happen when max_hits_per_query is honest
        s_row[found] = col;
      }
      s_bonus[found] = a.hit_bonus[h];
      s_has[found] = 1;
      if (a.hit_score) { s_score[found] = __ldcg(a.hit_score + (h - a.hit_rowptr[0])); s_has[found] = 2; }     // already re-scored
    }
    __syncthreads();
  }
  const int n = min(nsel + s_extra, MC);
  const int nc = a.hit_score ? nsel : n;                    // candidates phase C still has to re-score
  KEMR_SEL_STAMP(4);

  // C (two register sets, roles alternate).  An item is two rows re-scored together with interleaved reduction
  //   chains: the T2I and T2T row of one candidate, or -- single gallery -- the rows of two candidates.
  {
    CanonRow<NP> ra[2], rb[2];
    const bool two = a.G > 1;
    const double wa = a.wq[0] ? a.wq[0][qi] : a.w[0], wb = a.wq[0] ? a.wq[1][qi] : a.w[1];
    const int step = two ? W : 2 * W;                       // candidates consumed per round of the CTA
    auto fetch = [&](int set, int c) {
      const size_t off = (size_t)s_row[c] * a.D;
      ra[set].load(a.gal[0] + off, a.D, lane);
      if (two) rb[set].load(a.gal[1] + off, a.D, lane);
      else rb[set].load(a.gal[0] + (size_t)s_row[min(c + W, nc - 1)] * a.D, a.D, lane);
    };
    auto score = [&](int set, int c) {
      double sa, sb;
      cq.dot2_lane0(ra[set], rb[set], a.D, lane, sa, sb);
      if (lane == 0) {
        if (two) {
          if (s_has[c] != 2) s_score[c] = canon_fuse(sa, sb, true, wa, wb, a.alpha, s_bonus[c], s_has[c] != 0);
        } else {
          if (s_has[c] != 2) s_score[c] = canon_fuse(sa, 0.0, false, wa, wb, a.alpha, s_bonus[c], s_has[c] != 0);
          if (c + W < nc && s_has[c + W] != 2)
            s_score[c + W] = canon_fuse(sb, 0.0, false, wa, wb, a.alpha, s_bonus[c + W], s_has[c + W] != 0);
        }
      }
    };
    int c = warp;
    if (c < nc) fetch(0, c);
    while (c < nc) {
      if (c + step < nc) fetch(1, c + step);
      score(0, c);
      c += step;
      if (c >= nc) break;
      if (c + step < nc) fetch(0, c + step);
      score(1, c);
      c += step;
    }
  }
  __syncthreads();

  KEMR_SEL_STAMP(5);
  // D (+ the fused all-gather: the same rows go to this rank's slot of every rank's exchange buffer)
  unsigned int epoch = 0;
  size_t peer_o = 0;
  if (a.n_peer > 0) {
    epoch = *a.peer_epoch;
    peer_o = (size_t)(((long long)(epoch & 1u) * a.peer_world + a.peer_rank) * a.peer_block) + (size_t)qi * a.k;
  }
  auto emit = [&](int r, double sc, int64_t gi) {
    const size_t o = (size_t)qi * a.k + r;
    a.out_score64[o] = sc;
    if (a.out_score32) a.out_score32[o] = (float)sc;
    a.out_idx[o] = gi;
  };
  // rank by counting, one warp per candidate with the lanes spread over the rivals (a thread per candidate walking
  // all rivals is a chain of ~n dependent shared-memory compares: 2.7 us for 33 candidates at batch 1)
  for (int c = warp; c < n; c += W) {
    const double sc = s_score[c];
    const int32_t rc = s_row[c];
    int r = 0;
    for (int j = lane; j < n; j += 32) r += ahead64(s_score[j], s_row[j], sc, rc) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0 && r < a.k) {
      emit(r, sc, a.idx_base + rc);
      if (r == a.k - 1) s_kth = sc;
    }
  }
  for (int r = n + tid; r < a.k; r += T) emit(r, -INFINITY, -1);
  if (a.n_peer > 0) {
    // The fused all-gather: the k finished rows of this query go to this rank's slot of EVERY rank's exchange buffer.
    // The rows were just written one at a time by whichever lane ranked them; here the whole CTA copies them out with
    // consecutive threads on consecutive entries, so a peer receives each array as a few full NVLink write packets
    // instead of k scattered 8-byte stores issued by one lane (top-100, 4096 queries, 2 GPUs: the selection with the
    // push took 0.70 ms against 0.53 ms for selection + NCCL all-gather).
    __syncthreads();                                           // the rows above are visible to the whole CTA
    const double* ls = a.out_score64 + (size_t)qi * a.k;
    const int64_t* li = a.out_idx + (size_t)qi * a.k;
    for (int i = tid; i < a.k * a.n_peer; i += T) {
      const int d = i / a.k, r = i - d * a.k;
      reinterpret_cast<double*>(a.peer_base[d])[peer_o + r] = __ldcg(ls + r);          // L2: written by other threads of this CTA
      reinterpret_cast<int64_t*>(a.peer_base[d] + a.peer_idx_region)[peer_o + r] = (int64_t)__ldcg(reinterpret_cast<const long long*>(li) + r);
    }
    __threadfence_system();                                    // rows before the flag, at system scope (peer GPUs)
  }
  __syncthreads();
  if (a.n_peer > 0 && tid < a.n_peer) {
    unsigned int* flag = reinterpret_cast<unsigned int*>(a.peer_base[tid] + a.peer_flag_region) +
                         ((size_t)(epoch & 1u) * a.peer_world + a.peer_rank) * a.peer_max_q + qi;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
  }

  KEMR_SEL_STAMP(6);
  // E: nothing the scan or the stages above rejected can reach the k-th canonical score
  if (tid == 0) {
    int flag = 0;
    if (s_bound) {
      const double b = (double)key_score((uint64_t)s_bound) + a.eps * (1.0 + 1.0 / 64.0);
      const double reach = a.alpha * b + 1e-300;
      if (!(n >= a.k && s_kth > reach)) flag = 1;
    }
    a.out_flags[qi] = flag;
  }
}

// (The body keeps the query widened to binary64 plus two sets of candidate rows in registers: 128 per thread, four
// 4-warp CTAs per SM, so 1000 queries take two waves.  Measured and rejected: a register cap for 6 / 8 CTAs per SM --
// the spills cost more than the extra resident warps bring: C2 selection 26.8 -> 57.8 / 26.6 us, C1 56.3 -> 53.5 /
// 75.4 us, profiles/r02_session_m_stdout.txt.)
template <int NP, int W>
__global__ void __launch_bounds__(W * 32) select_kernel(SelectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // launched behind the scan with programmatic stream serialisation: wait here until the scan grid has completed and
  // its candidate lists are visible (returns at once for an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  select_query<NP, W>(a, blockIdx.x, smem_raw);
}

// ------------------------------------------------------------------ canonical pair scores
__global__ void score_pairs_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ ga,
                                   const uint16_t* __restrict__ gb, int64_t M, int D, double wa, double wb,
                                   const double* __restrict__ wqa, const double* __restrict__ wqb,
                                   double alpha, const int32_t* __restrict__ pq,
                                   const int64_t* __restrict__ prow, const double* __restrict__ pbonus,
                                   int64_t n, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = wid; i < n; i += nw) {
    const uint16_t* qrow = q + (size_t)pq[i] * D;
    if (prow[i] < 0 || prow[i] >= M) {                      // row outside the shard: no score (NaN ranks last)
      if (lane == 0) out[i] = __longlong_as_double(0x7ff8000000000000ll);
      continue;
    }
    const size_t off = (size_t)prow[i] * D;
    const double sa = canon_dot_warp(qrow, ga + off, D, lane);
    const double sb = gb ? canon_dot_warp(qrow, gb + off, D, lane) : 0.0;
    if (lane == 0)
      out[i] = canon_fuse(sa, sb, gb != nullptr, wqa ? wqa[pq[i]] : wa, wqa ? wqb[pq[i]] : wb, alpha, pbonus ? pbonus[i] : 0.0,
                          pbonus != nullptr);
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

// part_count is [P][q_stride] (q_stride = query rows padded by the scan kernel); outputs hold Q entries
__global__ void rank_sum_parts_kernel(const int32_t* __restrict__ part_count, int P, int Q, int q_stride,
                                      const unsigned int* __restrict__ amb_counter, unsigned int amb_cap,
                                      unsigned long long* __restrict__ count, int32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  unsigned long long s = 0;
  for (int p = 0; p < P; ++p) s += (unsigned long long)part_count[(size_t)p * q_stride + i];
  count[i] = s;
  flags[i] = (*amb_counter > amb_cap) ? 2 : 0;
}

struct RankFixArgs {
  const uint16_t* q;
  const uint16_t* gal[2];
  int G, D;
  double w[2];
  const double* wq[2];        // optional per-query weights [Q]
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
      const double wa = a.wq[0] ? a.wq[0][qi] : a.w[0], wb = a.wq[0] ? a.wq[1][qi] : a.w[1];
      const double f = canon_fuse(sa, sb, a.G > 1, wa, wb, a.alpha, 0.0, false);
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
        const double wa = a.wq[0] ? a.wq[0][qi] : a.w[0], wb = a.wq[0] ? a.wq[1][qi] : a.w[1];
        const double f0 = canon_fuse(sa, sb, a.G > 1, wa, wb, a.alpha, 0.0, false);
        const double f1 = canon_fuse(sa, sb, a.G > 1, wa, wb, a.alpha, bonus[h], true);
        const int d = (int)ahead64(f1, a.idx_base + row, a.t[qi], a.t_gidx[qi]) -
                      (int)ahead64(f0, a.idx_base + row, a.t[qi], a.t_gidx[qi]);
        if (d) atomicAdd(&a.count[qi], (unsigned long long)(long long)d);
      }
    }
  }
}

// ------------------------------------------------------------------ post-all-gather merge
// R per-shard lists of k entries, each already ordered by (score desc, idx asc) with empty slots (idx < 0) at the
// end.  The global position of an entry is its position in its own list plus, for every other list, the number of
// entries ahead of it -- a binary search, since the lists are sorted: R*log2(k) steps per entry instead of R*k.
// With `flags` (exchange over peer memory) list r of query qi is complete once flags[r * flag_stride + qi] holds the
// step's epoch: one thread per list waits for it (bounded: a rank that never arrives traps after ~20 s instead of
// hanging the GPU) before the lists are read.
__device__ __forceinline__ void merge_lists(const double* in_s, const int64_t* in_i, long long rank_stride,
                                            int R, int k, double* __restrict__ out_s,
                                            int64_t* __restrict__ out_i, const unsigned int* flags, long long flag_stride,
                                            unsigned int epoch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s = reinterpret_cast<double*>(smem_raw);
  int64_t* ix = reinterpret_cast<int64_t*>(s + (size_t)R * k);
  __shared__ int s_len[64];                 // valid entries per list (R <= 64)
  __shared__ int s_valid;
  const int qi = blockIdx.x;
  const int n = R * k;
  if (threadIdx.x < R) s_len[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_valid = 0;
  if (flags && threadIdx.x < R) {
    const unsigned int* f = flags + (size_t)threadIdx.x * flag_stride + qi;
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == epoch) break;
      if ((spins & 0xfffu) == 0xfffu) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (!t0) t0 = t;
        else if (t - t0 > 20000000000ull) __trap();
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const int r = c / k, j = c % k;
    const size_t o = (size_t)r * rank_stride + (size_t)qi * k + j;
    s[c] = __ldcg(in_s + o);                 // L2 is the point of coherence for rows a peer GPU stored
    ix[c] = (int64_t)__ldcg(reinterpret_cast<const long long*>(in_i) + o);
    if (ix[c] >= 0) { atomicAdd(&s_len[r], 1); atomicAdd(&s_valid, 1); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    if (ix[c] < 0) continue;
    const int r = c / k;
    int pos = c - r * k;                     // entries of its own list ahead of it
    for (int r2 = 0; r2 < R; ++r2) {
      if (r2 == r) continue;
      const double* ls = s + (size_t)r2 * k;
      const int64_t* li = ix + (size_t)r2 * k;
      int lo = 0, hi = s_len[r2];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ahead64(ls[mid], li[mid], s[c], ix[c])) lo = mid + 1; else hi = mid;
      }
      pos += lo;
    }
    if (pos < k) { out_s[(size_t)qi * k + pos] = s[c]; out_i[(size_t)qi * k + pos] = ix[c]; }
  }
  for (int r = s_valid + threadIdx.x; r < k; r += blockDim.x) {
    out_s[(size_t)qi * k + r] = -INFINITY;
    out_i[(size_t)qi * k + r] = -1;
  }
}

__global__ void merge_topk_kernel(const double* in_s, const int64_t* in_i, long long rank_stride, int R, int Q, int k,
                                  double* __restrict__ out_s, int64_t* __restrict__ out_i, const unsigned int* flags,
                                  long long flag_stride, const unsigned int* epoch_ptr) {
  (void)Q;
  merge_lists(in_s, in_i, rank_stride, R, k, out_s, out_i, flags, flag_stride, flags ? *epoch_ptr : 0u);
}

// plain gather out of this rank's exchange buffer: waits for rank r's flag of query qi, then copies its k rows to
// out[r][qi][:] (query-sharded replicas: every rank ends with every rank's results, no merge)
__global__ void gather_peer_kernel(const unsigned char* base, long long idx_region, long long flag_region, long long block,
                                   int world, int max_q, int Q, int k, double* __restrict__ out_s,
                                   int64_t* __restrict__ out_i, const unsigned int* epoch_ptr) {
  const int qi = blockIdx.x, r = blockIdx.y;
  const unsigned int epoch = *epoch_ptr;
  const size_t half = (size_t)(epoch & 1u) * world;
  if (threadIdx.x == 0) {
    const unsigned int* f = reinterpret_cast<const unsigned int*>(base + flag_region) + (half + r) * max_q + qi;
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == epoch) break;
      if ((spins & 0xfffu) == 0xfffu) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (!t0) t0 = t;
        else if (t - t0 > 20000000000ull) __trap();
      }
    }
  }
  __syncthreads();
  const double* ss = reinterpret_cast<const double*>(base) + (half + r) * block + (size_t)qi * k;
  const long long* si = reinterpret_cast<const long long*>(base + idx_region) + (half + r) * block + (size_t)qi * k;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    out_s[((size_t)r * Q + qi) * k + j] = __ldcg(ss + j);
    out_i[((size_t)r * Q + qi) * k + j] = (int64_t)__ldcg(si + j);
  }
}

// merge of the lists the ranks pushed into this rank's exchange buffer (kemr_peer_*): the epoch picks the buffer half
__global__ void merge_peer_kernel(const unsigned char* base, long long idx_region, long long flag_region, long long block,
                                  int world, int max_q, int Q, int k, double* __restrict__ out_s,
                                  int64_t* __restrict__ out_i, const unsigned int* epoch_ptr) {
  (void)Q;
  const unsigned int epoch = *epoch_ptr;
  const size_t half = (size_t)(epoch & 1u) * world;
  merge_lists(reinterpret_cast<const double*>(base) + half * block,
              reinterpret_cast<const int64_t*>(base + idx_region) + half * block, block, world, k, out_s, out_i,
              reinterpret_cast<const unsigned int*>(base + flag_region) + half * max_q, max_q, epoch);
}

}  // namespace kemr
