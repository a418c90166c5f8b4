// Matrix-taking compatibility kernels, the embedding boundary, the synthetic generator and the
// fused Recall@K / MRR / Mean-Rank reduction.  All HBM-bound streaming kernels.
#pragma once
#include "common.cuh"

namespace kemr {

// ------------------------------------------------------------------ fp32 rows -> bf16
// Replaces the normalise step of evaluator.py:120-135 plus the quantisation into the store.
__global__ void quantize_rows_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                     int64_t rows, int D, int normalize) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = wid; r < rows; r += nw) {
    const float* s = src + (size_t)r * D;
    float scale = 1.f;
    if (normalize) {
      float ss = 0.f;
      for (int d = lane; d < D; d += 32) { const float v = s[d]; ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
      scale = 1.f / sqrtf(ss);
    }
    uint16_t* o = dst + (size_t)r * D;
    for (int d = lane; d < D; d += 32) {
      const float v = normalize ? s[d] * scale : s[d];
      o[d] = f32_to_bf16_rne(v);
    }
  }
}

// Round-only variant for D % 128 == 0: every lane fetches all of its 16-byte pieces of the row up front (one round
// trip even when `src` is page-locked HOST memory read over PCIe -- the host-buffer search reads the caller's query
// array in place) and packs 4 bf16 per 8-byte store.  Element-wise identical to quantize_rows_kernel(normalize = 0).
template <int NV>   // float4 pieces per lane = D / 128
__global__ void quantize_rows_vec_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int D = NV * 128;
  for (int64_t r = wid; r < rows; r += nw) {
    const float4* s = reinterpret_cast<const float4*>(src + (size_t)r * D);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = s[lane + 32 * j];
    uint2* o = reinterpret_cast<uint2*>(dst + (size_t)r * D);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t lo = (uint32_t)f32_to_bf16_rne(v[j].x) | ((uint32_t)f32_to_bf16_rne(v[j].y) << 16);
      const uint32_t hi = (uint32_t)f32_to_bf16_rne(v[j].z) | ((uint32_t)f32_to_bf16_rne(v[j].w) << 16);
      o[lane + 32 * j] = make_uint2(lo, hi);
    }
  }
}

// ------------------------------------------------------------------ largest row norm of a bf16 matrix
// The selection margin eps of the scan is proportional to max||q|| * max||g||: one streaming
// pass, fp32 sum of squares per row (one warp per row), atomicMax on the bit pattern (norms are non-negative).
__global__ void row_norm_max_kernel(const uint16_t* __restrict__ x, int64_t rows, int D, float* __restrict__ out_max) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nchunk = D >> 3;
  float best = 0.f;
  for (int64_t r = wid; r < rows; r += nw) {
    const uint16_t* row = x + (size_t)r * D;
    float ss = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      const uint4 v = ldg_stream(row + (size_t)c * 8);
      const float a0 = bf16_lo(v.x), a1 = bf16_hi(v.x), a2 = bf16_lo(v.y), a3 = bf16_hi(v.y);
      const float a4 = bf16_lo(v.z), a5 = bf16_hi(v.z), a6 = bf16_lo(v.w), a7 = bf16_hi(v.w);
      ss = fmaf(a0, a0, ss); ss = fmaf(a1, a1, ss); ss = fmaf(a2, a2, ss); ss = fmaf(a3, a3, ss);
      ss = fmaf(a4, a4, ss); ss = fmaf(a5, a5, ss); ss = fmaf(a6, a6, ss); ss = fmaf(a7, a7, ss);
    }
    ss = warp_sum(ss);
    best = fmaxf(best, ss);
  }
  if (lane == 0) {
    const float n = sqrtf(best) * (1.0f + 1e-6f);                 // fp32 summation slack: never under-report
    if (isnan(n)) atomicMax(reinterpret_cast<unsigned int*>(out_max), 0x7f800000u);     // NaN rows: infinite margin
    else atomicMax(reinterpret_cast<unsigned int*>(out_max), __float_as_uint(n));
  }
}

// ------------------------------------------------------------------ synthetic gallery rows
__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ float synth_normal(uint64_t seed, int64_t grow, int d) {
  const uint64_t h = mix64(mix64(seed ^ 0x9e3779b97f4a7c15ull) + (uint64_t)grow * 0x100000001b3ull + (uint64_t)d);
  const float u1 = ((float)(uint32_t)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  const float u2 = ((float)(uint32_t)((h >> 16) & 0xffffff) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__global__ void synth_rows_kernel(uint16_t* __restrict__ dst, int64_t rows, int D, uint64_t seed,
                                  int64_t row_base) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = wid; r < rows; r += nw) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = synth_normal(seed, row_base + r, d); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float scale = 1.f / sqrtf(ss);
    uint16_t* o = dst + (size_t)r * D;
    for (int d = lane; d < D; d += 32) o[d] = f32_to_bf16_rne(synth_normal(seed, row_base + r, d) * scale);
  }
}

// ------------------------------------------------------------------ rank of a column in a given matrix
// metrics.py:34,62,68 on a caller-supplied matrix: stable descending order, NaN last.
// A target column outside [0, M) (a matrix with more rows than columns, metrics.py:37) has no match in the
// reference: `any(top_k == i)` is false and `argmax(all False) + 1` is 1.  It is reported as rank 0, which the
// reductions below read as "recall miss, position 1".
__global__ void __launch_bounds__(256) matrix_rank_kernel(const float* __restrict__ S, int Q, int64_t M,
                                                         int64_t ld, const int64_t* __restrict__ tcol,
                                                         int64_t* __restrict__ out_rank) {
  const int qi = blockIdx.x;
  const float* row = S + (size_t)qi * ld;
  const int64_t tc = tcol[qi];
  if (tc < 0 || tc >= M) {
    if (threadIdx.x == 0) out_rank[qi] = 0;
    return;
  }
  const float t = row[tc];
  const bool tnan = isnan(t);
  unsigned long long cnt = 0;
  for (int64_t j = threadIdx.x; j < M; j += blockDim.x) {
    const float v = row[j];
    const bool vnan = isnan(v);
    bool ahead;
    if (tnan) ahead = !vnan || (j < tc);
    else ahead = (v > t) || (v == t && j < tc);
    cnt += ahead ? 1ull : 0ull;
  }
  __shared__ unsigned long long s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) out_rank[qi] = (int64_t)s_cnt + 1;
}

// top-k columns of each row of a given matrix by (value desc, column asc)
__global__ void __launch_bounds__(256) matrix_topk_kernel(const float* __restrict__ S, int Q, int64_t M,
                                                         int64_t ld, int k, int64_t* __restrict__ out_idx,
                                                         float* __restrict__ out_val) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw);        // [8][k]
  const int qi = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = S + (size_t)qi * ld;
  uint64_t* mine = lists + (size_t)warp * k;
  for (int i = lane; i < k; i += 32) mine[i] = 0;
  __syncwarp();
  uint64_t thr = 0;
  for (int64_t j0 = (int64_t)warp * 32; j0 < M; j0 += 8 * 32) {
    const int64_t j = j0 + lane;
    uint64_t key = 0;
    if (j < M) key = make_key(row[j], (uint32_t)j);
    unsigned m = __ballot_sync(0xffffffffu, key > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t x = __shfl_sync(0xffffffffu, key, src);
      if (x > thr) { warp_list_insert(mine, k, x, lane); thr = mine[k - 1]; }
    }
  }
  __syncthreads();
  if (warp == 0) {
    thr = mine[k - 1];
    for (int w2 = 1; w2 < 8; ++w2) {
      const uint64_t* other = lists + (size_t)w2 * k;
      for (int i = 0; i < k; ++i) {
        const uint64_t x = other[i];
        if (x <= thr) break;
        warp_list_insert(mine, k, x, lane);
        thr = mine[k - 1];
      }
    }
    for (int i = lane; i < k; i += 32) {
      const uint64_t x = mine[i];
      out_idx[(size_t)qi * k + i] = x ? (int64_t)key_row(x) : -1;
      out_val[(size_t)qi * k + i] = x ? row[key_row(x)] : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------ dense KG fusion (bit-identical to numpy fp32)
__global__ void matrix_scale_kernel(const float* __restrict__ S, float* __restrict__ out, int Q, int64_t M,
                                    int64_t ld, int scale_first, float alpha32) {
  const int64_t total = (int64_t)Q * M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / M, c = i - r * M;
    const float v = S[(size_t)r * ld + c];
    // fusion.py:83  alpha*S + w*I  ->  non-hit entries see "+ 0.0"
    out[(size_t)r * ld + c] = scale_first ? __fadd_rn(__fmul_rn(alpha32, v), 0.0f) : v;
  }
}
// one thread per query row applies that row's hits in list order (duplicates add again,
// fusion.py:130), each add rounded to fp32 like numpy's in-place `+=` on a float32 array
__global__ void matrix_hits_kernel(float* __restrict__ out, int Q, int64_t M, int64_t ld,
                                   const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   const float* __restrict__ add) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= Q) return;
  for (int64_t h = rowptr[qi]; h < rowptr[qi + 1]; ++h) {
    const int64_t c = col[h];
    if (c < 0 || c >= M) continue;
    float* p = out + (size_t)qi * ld + c;
    *p = __fadd_rn(*p, add[h]);
  }
}

// ------------------------------------------------------------------ learned 2 -> H -> 1 fusion of two score matrices
// LinearFusionHead of the reference (fusion_model.py:25-48): out = w2 . relu(W1 [a, b] + b1) + b2 for every element,
// a = T2I score, b = T2T score, fp32 like torch.  W1 is [H][2] (nn.Linear layout), hidden units summed in
// increasing order.  Streams both matrices once; the H-unit MLP runs out of shared memory.
__global__ void __launch_bounds__(256) matrix_mlp2_kernel(const float* __restrict__ Sa, const float* __restrict__ Sb,
                                                         float* __restrict__ out, int64_t total,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, float b2, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* p = reinterpret_cast<float4*>(smem_raw);            // (w1[k][0], w1[k][1], b1[k], w2[k])
  for (int k = threadIdx.x; k < H; k += blockDim.x) p[k] = make_float4(w1[2 * k], w1[2 * k + 1], b1[k], w2[k]);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = Sa[i], b = Sb[i];
    float acc = 0.f;
    for (int k = 0; k < H; ++k) {
      const float4 c = p[k];
      const float h = fmaf(b, c.y, fmaf(a, c.x, c.z));
      acc = fmaf(fmaxf(h, 0.f), c.w, acc);
    }
    out[i] = acc + b2;
  }
}

// ------------------------------------------------------------------ fused InfoNCE rows (train/losses.py:45-55)
// row_loss[i] = logsumexp_j(a_i . b_j / tau) - a_i . b_i / tau, i.e. F.cross_entropy(logits, arange(B), reduction='none')
// of logits = A B^T / tau, without the (B, B) logits ever leaving the SM: one warp per row keeps a_i in registers and
// walks the rows of B (L2-resident: B x D fp32) four at a time -- per-lane partial dots, one transposed shuffle
// reduction for the four sums, online max / sum-of-exp per lane group.  fp32 like torch.  The symmetric loss is two
// calls, (A, B) and (B, A).
template <int CH>      // D <= 128 * CH: float4 pieces per lane
__global__ void __launch_bounds__(256) infonce_rows_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int B,
                                                          int D, float inv_tau, float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= B) return;
  const int nv = D >> 2;                                         // float4 pieces per row (D % 4 == 0)
  float4 a[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int p = lane + 32 * c;
    a[c] = p < nv ? reinterpret_cast<const float4*>(A + (size_t)i * D)[p] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // lane group g = lane >> 3 owns columns j with j % 4 == g (after the transposed reduction below)
  float m = -INFINITY, ssum = 0.f, diag = 0.f;
  for (int j0 = 0; j0 < B; j0 += 4) {
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = j0 + t;
      float acc = 0.f;
      if (j < B) {
        const float4* brow = reinterpret_cast<const float4*>(Bm + (size_t)j * D);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int p = lane + 32 * c;
          if (p < nv) {
            const float4 x = brow[p];
            acc = fmaf(a[c].x, x.x, acc); acc = fmaf(a[c].y, x.y, acc);
            acc = fmaf(a[c].z, x.z, acc); acc = fmaf(a[c].w, x.w, acc);
          }
        }
      }
      v[t] = acc;
    }
    // 4 sums over 32 lanes: halve the payload twice (steps 16, 8), then plain butterfly (4, 2, 1)
    {
      const bool up16 = (lane & 16) != 0;
      const float s0 = up16 ? v[0] : v[2], k0 = up16 ? v[2] : v[0];
      const float s1 = up16 ? v[1] : v[3], k1 = up16 ? v[3] : v[1];
      v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
      v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
      const bool up8 = (lane & 8) != 0;
      const float s2 = up8 ? v[0] : v[1], k2 = up8 ? v[1] : v[0];
      v[0] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    }
    const int t = ((lane & 16) ? 2 : 0) + ((lane & 8) ? 1 : 0);   // which of the four columns this lane now holds
    const int j = j0 + t;
    if (j < B) {
      const float logit = v[0] * inv_tau;
      if (j == i) diag = logit;
      const float mn = fmaxf(m, logit);
      ssum = ssum * __expf(m - mn) + __expf(logit - mn);          // m = -inf on the first column: exp(-inf) = 0
      m = mn;
    }
  }
  // merge the four lane groups' (max, sum) states and pick up the diagonal
#pragma unroll
  for (int o = 16; o >= 8; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, ssum, o);
    const float d2 = __shfl_xor_sync(0xffffffffu, diag, o);
    const float mn = fmaxf(m, m2);
    const float sa = m == -INFINITY ? 0.f : ssum * __expf(m - mn), sb = m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn);
    ssum = sa + sb; m = mn; diag += d2;
  }
  if (lane == 0) row_loss[i] = (m + logf(ssum)) - diag;
}

// ------------------------------------------------------------------ metrics reduction
// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum) over
// a[i] = 1.0 / rank[i]; reproduced so that MRR is bit-identical to np.mean(1.0/pos)*100.
// rank 0 = "target column not in the matrix": position 1 for MRR / Mean_Rank, never a recall hit (matrix_rank_kernel)
__host__ __device__ inline int64_t rank_pos(int64_t r) { return r > 0 ? r : 1; }
__host__ __device__ inline double recip_rank(const int64_t* r, int64_t i) { return 1.0 / (double)rank_pos(r[i]); }

__host__ __device__ inline double pairwise_leaf(const int64_t* r, int64_t off, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res += recip_rank(r, off + i);
    return res;
  }
  double acc[8];
  for (int j = 0; j < 8; ++j) acc[j] = recip_rank(r, off + j);
  int64_t i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) acc[j] += recip_rank(r, off + i + j);
  double res = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  for (; i < n; ++i) res += recip_rank(r, off + i);
  return res;
}

// explicit-stack traversal of numpy's split rule; leaf sums come from `leaf(off, n)`
template <class LeafFn>
__host__ __device__ inline double pairwise_tree(int64_t n_total, LeafFn leaf) {
  // post-order evaluation with a small stack: state 0 = visit, 1 = left done, 2 = right done
  struct Frame { int64_t off, n; double left; int state; };
  Frame st[64];
  int sp = 0;
  st[0] = Frame{0, n_total, 0.0, 0};
  double ret = 0.0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= 128) { ret = leaf(f.off, f.n); --sp; continue; }
    int64_t n2 = f.n / 2; n2 -= n2 % 8;
    if (f.state == 0) { f.state = 1; st[sp + 1] = Frame{f.off, n2, 0.0, 0}; ++sp; }
    else if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = Frame{f.off + n2, f.n - n2, 0.0, 0}; ++sp; }
    else { ret = f.left + ret; --sp; }
  }
  return ret;
}

constexpr int kMetricsThreads = 1024;
constexpr int kMaxLeaves = 8192;          // leaves hold 65..128 ranks each -> Q up to ~500k

__global__ void __launch_bounds__(kMetricsThreads) metrics_reduce_kernel(
    const int64_t* __restrict__ ranks, int Q, const int32_t* __restrict__ kv, int nk,
    int64_t* __restrict__ out_hits, double* __restrict__ out_stats, double* __restrict__ leaf_sums,
    int64_t* __restrict__ leaf_off, int64_t* __restrict__ leaf_n) {
  __shared__ unsigned long long s_hits[32];
  __shared__ unsigned long long s_sum;
  __shared__ int s_nleaf;
  if (threadIdx.x < 32) s_hits[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s_sum = 0;
    // enumerate leaves in traversal order
    int nl = 0;
    pairwise_tree(Q, [&](int64_t off, int64_t n) { leaf_off[nl] = off; leaf_n[nl] = n; ++nl; return 0.0; });
    s_nleaf = nl;
  }
  __syncthreads();
  unsigned long long hits[32];
  for (int i = 0; i < nk; ++i) hits[i] = 0;
  unsigned long long sum = 0;
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    const int64_t r = ranks[i];
    sum += (unsigned long long)rank_pos(r);
    for (int j = 0; j < nk; ++j) hits[j] += (r > 0 && r <= kv[j]) ? 1ull : 0ull;
  }
  atomicAdd(&s_sum, sum);
  for (int j = 0; j < nk; ++j) if (hits[j]) atomicAdd(&s_hits[j], hits[j]);
  for (int l = threadIdx.x; l < s_nleaf; l += blockDim.x) leaf_sums[l] = pairwise_leaf(ranks, leaf_off[l], leaf_n[l]);
  __syncthreads();
  if (threadIdx.x == 0) {
    int cursor = 0;
    const double rr = 0.0 + pairwise_tree(Q, [&](int64_t, int64_t) { return leaf_sums[cursor++]; });
    out_stats[0] = (double)s_sum;
    out_stats[1] = rr;
    for (int j = 0; j < nk; ++j) out_hits[j] = (int64_t)s_hits[j];
  }
}

}  // namespace kemr
