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
constexpr int kCountMax = 1024;          // up to this many gathered keys are ranked by counting
constexpr int kMaxParts = 4096;          // part lists per query the head pass can hold

// dynamic smem (16-byte aligned pieces first):
//   heads[P] u64 | keys[max(K*Kp, kSelectWarps*K)] u64 |
//   cand_score[max_cand] f64 | cand_bonus[max_cand] f64 | cand_row[max_cand] i32 | cand_has[max_cand] u8
inline size_t select_key_slots(int Kp, int K) {
  const size_t gathered = (size_t)K * Kp, per_warp = (size_t)kSelectWarps * K;
  return gathered > per_warp ? gathered : per_warp;
}
inline size_t select_smem_bytes(int P, int Kp, int K, int max_cand, int G, int D) {
  (void)G; (void)D;
  size_t b = (size_t)P * 8 + select_key_slots(Kp, K) * 8 + (size_t)max_cand * (8 + 8 + 4);
  b += ((size_t)max_cand + 15) & ~(size_t)15;
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(kSelectThreads) select_rescore_kernel(SelectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K, Kp = a.Kp, P = a.P;
  uint64_t* heads = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* lists = heads + P;
  const size_t gathered = (size_t)K * Kp, per_warp = (size_t)kSelectWarps * K;
  double* cand_score = reinterpret_cast<double*>(lists + (gathered > per_warp ? gathered : per_warp));
  double* cand_bonus = cand_score + a.max_cand;
  int32_t* cand_row = reinterpret_cast<int32_t*>(cand_bonus + a.max_cand);
  unsigned char* cand_has = reinterpret_cast<unsigned char*>(cand_row + a.max_cand);
  __shared__ int s_nsel, s_extra, s_nlist;
  __shared__ uint64_t s_sel[kMaxKSel];
  __shared__ int s_listid[kMaxKSel];
  __shared__ unsigned long long s_bound;     // largest key any stage rejected (0 = nothing rejected)

  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // this lane's share of the query row, widened to binary64 once
  CanonQuery cq;
  cq.load(a.q + (size_t)qi * a.D, a.D, lane);

  // ---- A1. heads of all part lists; only the K lists with the largest heads can hold a top-K key
  for (int i = threadIdx.x; i < K; i += blockDim.x) { s_sel[i] = 0; s_listid[i] = -1; }
  if (threadIdx.x == 0) { s_bound = 0; s_nlist = 0; s_extra = 0; }
  for (int p = threadIdx.x; p < P; p += blockDim.x) heads[p] = a.part_keys[((size_t)p * a.Q + qi) * Kp];
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const uint64_t x = heads[p];
    if (!x) continue;
    int r = 0;
    for (int j = 0; j < P; ++j) r += heads[j] > x ? 1 : 0;
    if (r < K) { s_listid[r] = p; atomicAdd(&s_nlist, 1); }
    else atomicMax(&s_bound, (unsigned long long)x);       // whole list rejected: nothing in it beats its head
  }
  __syncthreads();
  const int nlist = s_nlist;                                // selected lists occupy s_listid[0..nlist)

  // ---- A2. the K best keys among the selected lists
  const int n2 = nlist * Kp;
  if (n2 <= kCountMax) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const int l = i / Kp, j = i - l * Kp;
      lists[i] = a.part_keys[((size_t)s_listid[l] * a.Q + qi) * Kp + j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const uint64_t x = lists[i];
      if (!x) continue;
      int r = 0;
      for (int j = 0; j < n2; ++j) r += lists[j] > x ? 1 : 0;
      if (r < K) s_sel[r] = x;                              // keys are distinct -> ranks are distinct
      // rejected here, or last entry of a full part list (that part rejected rows below it)
      if (r >= K || (i % Kp) == Kp - 1) atomicMax(&s_bound, (unsigned long long)x);
    }
    __syncthreads();
  } else {
    uint64_t* mine = lists + (size_t)warp * K;
    for (int i = lane; i < K; i += 32) mine[i] = 0;
    __syncwarp();
    uint64_t thr = 0;
    unsigned long long bnd = 0;
    for (int l = warp; l < nlist; l += kSelectWarps) {
      const uint64_t* src = a.part_keys + ((size_t)s_listid[l] * a.Q + qi) * Kp;
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
          if (x <= thr) { if (lane == 0 && x) atomicMax(&s_bound, (unsigned long long)x); break; }
          warp_list_insert(mine, K, x, lane);
          thr = mine[K - 1];
        }
      }
      // keys displaced from the final list are bounded by its last key
      if (lane == 0 && mine[K - 1]) atomicMax(&s_bound, (unsigned long long)mine[K - 1]);
      for (int i = lane; i < K; i += 32) s_sel[i] = mine[i];
    }
    __syncthreads();
  }
  if (warp == 0) {
    int n = 0;
    for (int i = lane; i < K; i += 32) n += (s_sel[i] != 0);
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) s_nsel = n;
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

  // ---- C. canonical re-scoring, one warp per candidate, straight from coalesced 16-byte loads
  for (int c = warp; c < n; c += kSelectWarps) {
    const size_t off = (size_t)cand_row[c] * a.D;
    const double sa = canon_dot_q(cq, a.gal[0] + off, a.D, lane);
    const double sb = a.G > 1 ? canon_dot_q(cq, a.gal[1] + off, a.D, lane) : 0.0;
    if (lane == 0)
      cand_score[c] = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, cand_bonus[c], cand_has[c] != 0);
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

// ------------------------------------------------------------------ fast path: one WARP per query
// Same contract as select_rescore_kernel for the common small case (few parts, short lists, at
// most 64 candidates): no block barriers, 4 independent queries per CTA, so a 1000-query batch
// fits the GPU in a single wave and the re-scoring loads of different queries overlap.
constexpr int kSelWarpWarps = 4;
constexpr int kSelWarpMaxP = 64;
constexpr int kSelWarpMaxKeys = 256;
constexpr int kSelWarpMaxCand = 64;

inline bool select_warp_ok(int P, int Kp, int K, int max_cand) {
  const int nl = P < K ? P : K;
  return P <= kSelWarpMaxP && K <= 32 && nl * Kp <= kSelWarpMaxKeys && max_cand <= kSelWarpMaxCand;
}

template <int NP>
__global__ void __launch_bounds__(kSelWarpWarps * 32) select_warp_kernel(SelectArgs a, int nq) {
  __shared__ uint64_t s_heads[kSelWarpWarps][kSelWarpMaxP];
  __shared__ uint64_t s_keys[kSelWarpWarps][kSelWarpMaxKeys];
  __shared__ uint64_t s_selk[kSelWarpWarps][32];
  __shared__ int s_lid[kSelWarpWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.x * kSelWarpWarps + warp;
  if (qi >= nq) return;
  const int K = a.K, Kp = a.Kp, P = a.P;
  uint64_t* heads = s_heads[warp];
  uint64_t* keys = s_keys[warp];
  uint64_t* selk = s_selk[warp];
  int* lid = s_lid[warp];
  unsigned long long bound = 0;                    // per-lane partial, max-reduced at the end

  CanonQueryT<NP> cq;
  cq.load(a.q + (size_t)qi * a.D, a.D, lane);

  // A1. heads of the part lists -> the (at most K) lists that can hold a top-K key
  for (int p = lane; p < P; p += 32) heads[p] = a.part_keys[((size_t)p * a.Q + qi) * Kp];
  selk[lane] = 0; lid[lane] = -1;
  __syncwarp();
  int nlist = 0;
  for (int p = lane; p < P; p += 32) {
    const uint64_t x = heads[p];
    if (!x) continue;
    int r = 0;
    for (int j = 0; j < P; ++j) r += heads[j] > x ? 1 : 0;
    if (r < K) { lid[r] = p; ++nlist; }
    else if (x > bound) bound = x;
  }
  for (int o = 16; o > 0; o >>= 1) nlist += __shfl_xor_sync(0xffffffffu, nlist, o);
  __syncwarp();

  // A2. K best keys among the selected lists
  const int n2 = nlist * Kp;
  for (int i = lane; i < n2; i += 32) {
    const int l = i / Kp, j = i - l * Kp;
    keys[i] = a.part_keys[((size_t)lid[l] * a.Q + qi) * Kp + j];
  }
  __syncwarp();
  for (int i = lane; i < n2; i += 32) {
    const uint64_t x = keys[i];
    if (!x) continue;
    int r = 0;
    for (int j = 0; j < n2; ++j) r += keys[j] > x ? 1 : 0;
    if (r < K) selk[r] = x;
    if ((r >= K || (i % Kp) == Kp - 1) && x > bound) bound = x;
  }
  __syncwarp();
  const uint64_t mykey = selk[lane];                // lane i <-> candidate i (K <= 32)
  const int nsel = __popc(__ballot_sync(0xffffffffu, mykey != 0));

  // B. candidates: lane holds candidate `lane` and (KG hits) candidate `32+lane`
  int32_t row0 = mykey ? (int32_t)key_row(mykey) : -1, row1 = -1;
  double bon0 = 0.0, bon1 = 0.0;
  bool has0 = false, has1 = false;
  int n = nsel;
  if (a.hit_rowptr) {
    const int64_t h0 = a.hit_rowptr[qi], h1 = a.hit_rowptr[qi + 1];
    for (int64_t h = h0; h < h1; ++h) {             // few dozen hits: walk them warp-uniformly
      const int32_t col = a.hit_col[h];
      if (col < 0 || (int64_t)col >= a.M) continue;
      const double b = a.hit_bonus[h];
      const unsigned m0 = __ballot_sync(0xffffffffu, row0 == col);
      const unsigned m1 = __ballot_sync(0xffffffffu, row1 == col);
      if (m0) { if (row0 == col) { bon0 = b; has0 = true; } }
      else if (m1) { if (row1 == col) { bon1 = b; has1 = true; } }
      else if (n < kSelWarpMaxCand) {
        if (n < 32) { if (lane == n) { row0 = col; bon0 = b; has0 = true; } }
        else if (lane == n - 32) { row1 = col; bon1 = b; has1 = true; }
        ++n;
      }
    }
  }

  // C. canonical re-scoring; the whole warp works on one candidate at a time, and the rows of
  //    candidate c+1 (both galleries) are already in flight while candidate c is accumulated
  double sc0 = 0.0, sc1 = 0.0;
  CanonRow<NP> cur[2], nxt[2];
  if (n > 0) {
    const size_t off = (size_t)__shfl_sync(0xffffffffu, row0, 0) * a.D;
    cur[0].load(a.gal[0] + off, a.D, lane);
    if (a.G > 1) cur[1].load(a.gal[1] + off, a.D, lane);
  }
  for (int c = 0; c < n; ++c) {
    if (c + 1 < n) {
      const int32_t rown = __shfl_sync(0xffffffffu, c + 1 < 32 ? row0 : row1, (c + 1) & 31);
      const size_t off = (size_t)rown * a.D;
      nxt[0].load(a.gal[0] + off, a.D, lane);
      if (a.G > 1) nxt[1].load(a.gal[1] + off, a.D, lane);
    }
    const double sa = cq.dot(cur[0], a.D, lane);
    const double sb = a.G > 1 ? cq.dot(cur[1], a.D, lane) : 0.0;
    if (lane == (c & 31)) {
      if (c < 32) sc0 = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, bon0, has0);
      else sc1 = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, bon1, has1);
    }
    cur[0] = nxt[0];
    cur[1] = nxt[1];
  }

  // D. order by (score desc, row asc) by counting over shuffled copies; write the first k
  int r0 = 0, r1 = 0;
  for (int j = 0; j < n; ++j) {
    const double sj = __shfl_sync(0xffffffffu, j < 32 ? sc0 : sc1, j & 31);
    const int32_t rj = __shfl_sync(0xffffffffu, j < 32 ? row0 : row1, j & 31);
    r0 += ahead64(sj, rj, sc0, row0) ? 1 : 0;
    r1 += ahead64(sj, rj, sc1, row1) ? 1 : 0;
  }
  double kth = -INFINITY;
  if (lane < n && r0 < a.k) {
    const size_t o = (size_t)qi * a.k + r0;
    a.out_score64[o] = sc0;
    if (a.out_score32) a.out_score32[o] = (float)sc0;
    a.out_idx[o] = a.idx_base + row0;
    if (r0 == a.k - 1) kth = sc0;
  }
  if (32 + lane < n && r1 < a.k) {
    const size_t o = (size_t)qi * a.k + r1;
    a.out_score64[o] = sc1;
    if (a.out_score32) a.out_score32[o] = (float)sc1;
    a.out_idx[o] = a.idx_base + row1;
    if (r1 == a.k - 1) kth = sc1;
  }
  for (int r = n + lane; r < a.k; r += 32) {
    const size_t o = (size_t)qi * a.k + r;
    a.out_score64[o] = -INFINITY;
    if (a.out_score32) a.out_score32[o] = -INFINITY;
    a.out_idx[o] = -1;
  }

  // E. certificate
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bound, o);
    bound = ob > bound ? ob : bound;
    kth = fmax(kth, __shfl_xor_sync(0xffffffffu, kth, o));
  }
  if (lane == 0) {
    int flag = 0;
    if (bound) {
      const double b = (double)key_score((uint64_t)bound) + a.eps * (1.0 + 1.0 / 64.0);
      const double reach = a.alpha * b + 1e-300;
      if (!(n >= a.k && kth > reach)) flag = 1;
    }
    a.out_flags[qi] = flag;
  }
}

// ------------------------------------------------------------------ fast path: one small CTA per query
// Same contract and the same small-case limits as select_warp_kernel, but W warps share one query:
// the merge of the part lists is done by all W*32 threads, and the candidates are dealt round-robin
// to the warps, each keeping the rows of its NEXT candidate in flight while it accumulates the current
// one.  select_warp_kernel walks a query's 16+ candidates one after another in a single warp, so a
// 1000-query batch is one wave of 1000 warps whose duration is the length of that serial chain
// (~34 us on C2, ncu); here the chain is W times shorter and W times as many row loads are in flight.
template <int NP, int W>
__global__ void __launch_bounds__(W * 32) select_query_kernel(SelectArgs a) {
  constexpr int T = W * 32;
  __shared__ uint64_t s_heads[kSelWarpMaxP];
  __shared__ uint64_t s_keys[kSelWarpMaxKeys];
  __shared__ uint64_t s_selk[32];
  __shared__ int s_lid[32];
  __shared__ int32_t s_row[kSelWarpMaxCand];
  __shared__ double s_bonus[kSelWarpMaxCand];
  __shared__ double s_score[kSelWarpMaxCand];
  __shared__ unsigned char s_has[kSelWarpMaxCand];
  __shared__ unsigned long long s_bound;
  __shared__ int s_nlist, s_extra, s_nsurv;
  __shared__ double s_kth;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qi = blockIdx.x;
  const int K = a.K, Kp = a.Kp, P = a.P;

  CanonQueryT<NP> cq;
  cq.load(a.q + (size_t)qi * a.D, a.D, lane);

  // A1. heads of the part lists -> the (at most K) lists that can hold a top-K key
  for (int p = tid; p < P; p += T) s_heads[p] = a.part_keys[((size_t)p * a.Q + qi) * Kp];
  if (tid < 32) { s_selk[tid] = 0; s_lid[tid] = -1; }
  if (tid == 0) { s_bound = 0; s_nlist = 0; s_extra = 0; s_nsurv = 0; s_kth = -INFINITY; }
  __syncthreads();
  for (int p = tid; p < P; p += T) {
    const uint64_t x = s_heads[p];
    if (!x) continue;
    int r = 0;
    for (int j = 0; j < P; ++j) r += s_heads[j] > x ? 1 : 0;
    if (r < K) { s_lid[r] = p; atomicAdd(&s_nlist, 1); }
    else atomicMax(&s_bound, (unsigned long long)x);
  }
  __syncthreads();

  // A2. K best keys among the selected lists.  When K lists were selected their heads alone are K keys
  //     >= the K-th head, so anything below that head is out; only the survivors are ranked by counting.
  const int nlist = s_nlist;
  const uint64_t cut = nlist >= K ? s_heads[s_lid[K - 1]] : 0ull;
  {
    const int n2 = nlist * Kp;
    unsigned long long rej = 0;                          // largest key this thread saw rejected
    for (int i = tid; i < n2; i += T) {
      const int l = i / Kp, j = i - l * Kp;
      const uint64_t x = a.part_keys[((size_t)s_lid[l] * a.Q + qi) * Kp + j];
      if (!x) continue;
      if (x >= cut) s_keys[atomicAdd(&s_nsurv, 1)] = x;
      else if (x > rej) rej = x;
      if (j == Kp - 1 && x > rej) rej = x;               // a full part list rejected rows below its last key
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, rej, o);
      rej = other > rej ? other : rej;
    }
    if (lane == 0 && rej) atomicMax(&s_bound, rej);
  }
  __syncthreads();
  const int ns = s_nsurv;
  for (int i = tid; i < ns; i += T) {
    const uint64_t x = s_keys[i];
    int r = 0;
    for (int j = 0; j < ns; ++j) r += s_keys[j] > x ? 1 : 0;
    if (r < K) s_selk[r] = x;                            // keys are distinct -> ranks are distinct
    else atomicMax(&s_bound, (unsigned long long)x);
  }
  __syncthreads();
  int nsel = __popc(__ballot_sync(0xffffffffu, s_selk[lane] != 0));          // keys fill s_selk[0..nsel), descending
  // Candidates past the k-th whose fp32 score lies more than 2 eps below the k-th fp32 score cannot reach
  // the top k (|fp32 - canonical| <= eps; a KG bonus only raises the k-th score): they are not re-scored
  // and count as rejected for the certificate.  KG hits among them come back through the hit list.
  if (nsel > a.k) {
    const double kth32 = (double)key_score(s_selk[a.k - 1]) - 2.0 * a.eps * (1.0 + 1.0 / 64.0);
    const bool drop = lane >= a.k && lane < nsel && (double)key_score(s_selk[lane]) < kth32;
    const unsigned m = __ballot_sync(0xffffffffu, drop);
    if (m) {
      const int first = __ffs(m) - 1;                    // descending order: everything from `first` on drops
      if (tid == 0) atomicMax(&s_bound, (unsigned long long)s_selk[first]);
      nsel = first;
    }
  }

  // B. candidate table = scan candidates U KG hits
  if (tid < nsel) { s_row[tid] = (int32_t)key_row(s_selk[tid]); s_bonus[tid] = 0.0; s_has[tid] = 0; }
  __syncthreads();
  if (a.hit_rowptr) {
    const int64_t h0 = a.hit_rowptr[qi], h1 = a.hit_rowptr[qi + 1];
    for (int64_t h = h0 + tid; h < h1; h += T) {
      const int32_t col = a.hit_col[h];
      if (col < 0 || (int64_t)col >= a.M) continue;
      int found = -1;
      for (int i = 0; i < nsel; ++i) if (s_row[i] == col) { found = i; break; }
      if (found < 0) {
        found = nsel + atomicAdd(&s_extra, 1);
        if (found >= kSelWarpMaxCand) continue;      // cannot happen when max_hits_per_query is honest
        s_row[found] = col;
      }
      s_bonus[found] = a.hit_bonus[h];
      s_has[found] = 1;
    }
    __syncthreads();
  }
  const int n = min(nsel + s_extra, kSelWarpMaxCand);

  // C. canonical re-scoring: warp w takes candidates w, w+W, ...; the rows of its next candidate are in
  //    flight while the current one is accumulated (two register sets, roles alternate)
  {
    CanonRow<NP> ra[2], rb[2];
    auto fetch = [&](int set, int c) {
      const size_t off = (size_t)s_row[c] * a.D;
      ra[set].load(a.gal[0] + off, a.D, lane);
      if (a.G > 1) rb[set].load(a.gal[1] + off, a.D, lane);
    };
    auto score = [&](int set, int c) {
      double sa, sb = 0.0;
      if (a.G > 1) cq.dot2_lane0(ra[set], rb[set], a.D, lane, sa, sb);
      else sa = cq.dot(ra[set], a.D, lane);
      if (lane == 0) s_score[c] = canon_fuse(sa, sb, a.G > 1, a.w[0], a.w[1], a.alpha, s_bonus[c], s_has[c] != 0);
    };
    int c = warp;
    if (c < n) fetch(0, c);
    while (c < n) {
      if (c + W < n) fetch(1, c + W);
      score(0, c);
      c += W;
      if (c >= n) break;
      if (c + W < n) fetch(0, c + W);
      score(1, c);
      c += W;
    }
  }
  __syncthreads();

  // D. order by (score desc, row asc) by counting; write the first k
  for (int c = tid; c < n; c += T) {
    const double sc = s_score[c];
    const int32_t rc = s_row[c];
    int r = 0;
    for (int j = 0; j < n; ++j) r += ahead64(s_score[j], s_row[j], sc, rc) ? 1 : 0;
    if (r < a.k) {
      const size_t o = (size_t)qi * a.k + r;
      a.out_score64[o] = sc;
      if (a.out_score32) a.out_score32[o] = (float)sc;
      a.out_idx[o] = a.idx_base + rc;
      if (r == a.k - 1) s_kth = sc;
    }
  }
  for (int r = n + tid; r < a.k; r += T) {
    const size_t o = (size_t)qi * a.k + r;
    a.out_score64[o] = -INFINITY;
    if (a.out_score32) a.out_score32[o] = -INFINITY;
    a.out_idx[o] = -1;
  }
  __syncthreads();

  // E. certificate
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
