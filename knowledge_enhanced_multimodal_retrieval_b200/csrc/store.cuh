// Gallery-side data path either side of the scan: the KG-hit CSR builder on the device, the
// uuid -> row map and the persisted bf16 embedding store on the host.
//
// KG hits arrive per query as a LIST of gallery rows in the order the knowledge graph returned them (the reference
// walks the same lists in Python: fusion.py:68-80, :119-130, :180-204), already mapped to GLOBAL row ids (-1 =
// unknown artefact).  The scan wants, per gallery shard [lo, hi), a CSR with UNIQUE local columns per query and the
// bonus already aggregated: indicator semantics keep one listing (fusion.py:80), additive semantics add the bonus
// once per listing (fusion.py:130).  Lists are short (dozens of rows), so one warp handles one query and finds
// repeats by comparing against the earlier entries of the same list; output order = order of first occurrence.
#pragma once
#include "common.cuh"

namespace kemr {

constexpr int kHitsWarpsPerBlock = 8;

// For list entry i of query qi: is it inside the shard, and is it the first listing of its row?  Returns the number
// of listings of that row in the whole list (0 if this is not the first one / out of range).
__device__ __forceinline__ int hits_first_count(const int64_t* __restrict__ rows, int64_t n, int64_t i, int64_t lo,
                                                int64_t hi) {
  const int64_t r = rows[i];
  if (r < lo || r >= hi) return 0;
  for (int64_t j = 0; j < i; ++j)
    if (rows[j] == r) return 0;
  int c = 1;
  for (int64_t j = i + 1; j < n; ++j) c += rows[j] == r ? 1 : 0;
  return c;
}

// pass 1: unique in-shard rows per query
__global__ void hits_count_kernel(const int64_t* __restrict__ list_rowptr, const int64_t* __restrict__ list_rows, int Q,
                                  int64_t lo, int64_t hi, int64_t* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= Q) return;
  const int64_t b = list_rowptr[qi], n = list_rowptr[qi + 1] - b;
  int c = 0;
  for (int64_t i = lane; i < n; i += 32) c += hits_first_count(list_rows + b, n, i, lo, hi) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) out_count[qi] = c;
}

// exclusive scan of Q counts into rowptr[Q+1] (single CTA; Q is a query batch, not a gallery) + max per query
__global__ void __launch_bounds__(1024) hits_scan_kernel(const int64_t* __restrict__ count, int Q,
                                                         int64_t* __restrict__ rowptr, int64_t* __restrict__ out_max) {
  __shared__ int64_t s_part[1024];
  __shared__ int64_t s_max[1024];
  const int t = threadIdx.x;
  const int per = (Q + 1023) / 1024;
  const int b = t * per, e = min(Q, b + per);
  int64_t sum = 0, mx = 0;
  for (int i = b; i < e; ++i) { sum += count[i]; mx = max(mx, count[i]); }
  s_part[t] = sum; s_max[t] = mx;
  __syncthreads();
  if (t == 0) {
    int64_t run = 0, m = 0;
    for (int i = 0; i < 1024; ++i) { const int64_t v = s_part[i]; s_part[i] = run; run += v; m = max(m, s_max[i]); }
    rowptr[Q] = run;
    *out_max = m;
  }
  __syncthreads();
  int64_t run = s_part[t];
  for (int i = b; i < e; ++i) { rowptr[i] = run; run += count[i]; }
}

// pass 2: write (local column, aggregated bonus) in order of first occurrence
__global__ void hits_fill_kernel(const int64_t* __restrict__ list_rowptr, const int64_t* __restrict__ list_rows,
                                 const double* __restrict__ bonus_per_query, int Q, int64_t lo, int64_t hi, int sum_repeats,
                                 const int64_t* __restrict__ rowptr, int32_t* __restrict__ out_col,
                                 double* __restrict__ out_bonus) {
  const int lane = threadIdx.x & 31;
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= Q) return;
  const int64_t b = list_rowptr[qi], n = list_rowptr[qi + 1] - b;
  const double bonus = bonus_per_query[qi];
  int64_t w = rowptr[qi];
  for (int64_t i0 = 0; i0 < n; i0 += 32) {
    const int64_t i = i0 + lane;
    const int c = i < n ? hits_first_count(list_rows + b, n, i, lo, hi) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, c != 0);
    if (c) {
      const int64_t o = w + __popc(m & ((1u << lane) - 1u));
      out_col[o] = (int32_t)(list_rows[b + i] - lo);
      double v = bonus;
      if (sum_repeats) for (int r = 1; r < c; ++r) v = __dadd_rn(v, bonus);     // the reference adds once per listing
      out_bonus[o] = v;
    }
    w += __popc(m);
  }
}

// ------------------------------------------------------------------ CSR -> CSR: query subset and / or row shard
// out query i = in query sel[i] (sel == null: identity); entries with col in [lo, hi) survive, re-based to col - lo,
// order kept.  One warp per output query; used for the per-shard CSR of a row-sharded gallery (SURVEY section 8e:
// "split each query's hit list by shard") and for the re-run of uncertified queries.
__global__ void hits_filter_count_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                         const int64_t* __restrict__ sel, int Q_out, int64_t lo, int64_t hi,
                                         int64_t* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const int qo = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qo >= Q_out) return;
  const int64_t qi = sel ? sel[qo] : qo;
  const int64_t b = rowptr[qi], e = rowptr[qi + 1];
  int c = 0;
  for (int64_t i = b + lane; i < e; i += 32) { const int64_t x = col[i]; c += (x >= lo && x < hi) ? 1 : 0; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) out_count[qo] = c;
}
__global__ void hits_filter_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                        const double* __restrict__ bonus, const int64_t* __restrict__ sel, int Q_out,
                                        int64_t lo, int64_t hi, const int64_t* __restrict__ out_rowptr,
                                        int32_t* __restrict__ out_col, double* __restrict__ out_bonus) {
  const int lane = threadIdx.x & 31;
  const int qo = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qo >= Q_out) return;
  const int64_t qi = sel ? sel[qo] : qo;
  const int64_t b = rowptr[qi], e = rowptr[qi + 1];
  int64_t w = out_rowptr[qo];
  for (int64_t i0 = b; i0 < e; i0 += 32) {
    const int64_t i = i0 + lane;
    const int64_t x = i < e ? (int64_t)col[i] : -1;
    const bool keep = i < e && x >= lo && x < hi;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t o = w + __popc(m & ((1u << lane) - 1u));
      out_col[o] = (int32_t)(x - lo);
      out_bonus[o] = bonus[i];
    }
    w += __popc(m);
  }
}

// bonus of each query's own target column (0 when the target is not a KG hit of that query; columns are unique)
__global__ void hits_target_bonus_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                         const double* __restrict__ bonus, int Q, const int64_t* __restrict__ target,
                                         double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= Q) return;
  const int64_t t = target[qi];
  double v = 0.0;
  for (int64_t i = rowptr[qi] + lane; i < rowptr[qi + 1]; i += 32)
    if ((int64_t)col[i] == t) v = bonus[i];
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = v != 0.0 ? v : other;
  }
  if (lane == 0) out[qi] = v;
}

}  // namespace kemr
