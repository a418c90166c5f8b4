/*
 * CPU ORACLE (C restatement) of the canonical scoring contract.  TEST INFRASTRUCTURE ONLY:
 * built into oracle/_build/libkemr_oracle.so by oracle/Makefile and loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product package.
 *
 * It restates, for sizes numpy cannot cover in seconds, what oracle/oracle.py defines:
 *   score   = sum_d q[d]*g[d] in binary64: 256 running sums (d mod 256, increasing d), 8-to-1 tree
 *             per 16-byte piece, 32 pieces folded 16,8,4,2,1       (oracle.py::canon_dot64)
 *   clip    = fl(fl(w_a*S_a) + fl(w_b*S_b)); final = fl(fl(alpha*clip) + bonus)  (canon_fused64)
 *   ranking = final descending, ties by lowest index                 (canon_topk / canon_rank)
 * which in turn pins the reference's similarity + fusion + argsort path
 * (src/clip/eval/metrics.py:102,145-148,34,62; src/clip/eval/fusion.py:83).
 * Compile WITHOUT -ffast-math: the summation order is the contract.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double bf16_to_f64(uint16_t b) {
  uint32_t u = ((uint32_t)b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return (double)f;
}

/* canonical dot of one query (already widened to double) with one bf16 row: 256 running sums
 * (position inside a 256-element block, blocks in increasing order), 8-to-1 tree per 16-byte piece,
 * then the 32 pieces fold 16,8,4,2,1 */
static inline double canon_dot(const double* q, const uint16_t* g, int D) {
  double acc[256];
  for (int r = 0; r < 256; ++r) acc[r] = 0.0;
  for (int d0 = 0; d0 < D; d0 += 256) {
    const int n = D - d0 < 256 ? D - d0 : 256;
    for (int r = 0; r < n; ++r) acc[r] += q[d0 + r] * bf16_to_f64(g[d0 + r]);   /* product exact in binary64 */
  }
  double lane[32];
  for (int l = 0; l < 32; ++l) {
    const double* a = acc + 8 * l;
    lane[l] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  for (int off = 16; off >= 1; off >>= 1)
    for (int l = 0; l < off; ++l) lane[l] = lane[l] + lane[l + off];
  return lane[0];
}

static inline double fuse(double sa, double sb, int two, double wa, double wb, double alpha) {
  volatile double clip = wa * sa;                  /* volatile: forbid contraction into fma */
  if (two) { volatile double t = wb * sb; clip = clip + t; }
  volatile double f = alpha * clip;
  return f;
}

static inline int ahead(double sa, int64_t ia, double sb, int64_t ib) {
  const int na = isnan(sa), nb = isnan(sb);
  if (na || nb) return (!na && nb) || (na && nb && ia < ib);
  return sa > sb || (sa == sb && ia < ib);
}

/* optional per-query fusion weights [Q] (gated fusion heads, fusion_model.py:21,178,194); NULL = the scalars.
 * Set before a call with oracle_set_query_weights, cleared with (NULL, NULL). */
static const double* g_wq_a = 0;
static const double* g_wq_b = 0;
void oracle_set_query_weights(const double* wq_a, const double* wq_b) { g_wq_a = wq_a; g_wq_b = wq_b; }

/* final scores of query qi against all M rows -> out[M]; hits (CSR, unique cols) add their bonus */
static void score_row(const uint16_t* q, const uint16_t* ga, const uint16_t* gb, int64_t M, int D, double wa,
                      double wb, double alpha, const int64_t* rowptr, const int32_t* col, const double* bonus,
                      int qi, double* out, double* qd) {
  for (int d = 0; d < D; ++d) qd[d] = bf16_to_f64(q[(size_t)qi * D + d]);
  if (g_wq_a) { wa = g_wq_a[qi]; wb = g_wq_b[qi]; }
  for (int64_t j = 0; j < M; ++j) {
    const double sa = canon_dot(qd, ga + (size_t)j * D, D);
    const double sb = gb ? canon_dot(qd, gb + (size_t)j * D, D) : 0.0;
    out[j] = fuse(sa, sb, gb != 0, wa, wb, alpha);
  }
  if (rowptr)
    for (int64_t h = rowptr[qi]; h < rowptr[qi + 1]; ++h) {
      volatile double f = out[col[h]] + bonus[h];
      out[col[h]] = f;
    }
}

/* dense canonical final scores [Q, M] */
int oracle_scores(const uint16_t* q, int Q, const uint16_t* ga, const uint16_t* gb, int64_t M, int D, double wa,
                  double wb, double alpha, const int64_t* rowptr, const int32_t* col, const double* bonus,
                  double* out) {
#pragma omp parallel
  {
    double* qd = (double*)malloc(sizeof(double) * (size_t)D);
#pragma omp for schedule(dynamic, 1)
    for (int qi = 0; qi < Q; ++qi)
      score_row(q, ga, gb, M, D, wa, wb, alpha, rowptr, col, bonus, qi, out + (size_t)qi * M, qd);
    free(qd);
  }
  return 0;
}

/* top-k (score desc, index asc) and/or 1-based rank of target[qi]; any output may be NULL */
int oracle_topk_rank(const uint16_t* q, int Q, const uint16_t* ga, const uint16_t* gb, int64_t M, int D, double wa,
                     double wb, double alpha, const int64_t* rowptr, const int32_t* col, const double* bonus,
                     int k, int64_t* out_idx, double* out_score, const int64_t* target, int64_t* out_rank) {
#pragma omp parallel
  {
    double* qd = (double*)malloc(sizeof(double) * (size_t)D);
    double* s = (double*)malloc(sizeof(double) * (size_t)M);
#pragma omp for schedule(dynamic, 1)
    for (int qi = 0; qi < Q; ++qi) {
      score_row(q, ga, gb, M, D, wa, wb, alpha, rowptr, col, bonus, qi, s, qd);
      if (out_idx) {
        int64_t* bi = out_idx + (size_t)qi * k;
        double* bs = out_score + (size_t)qi * k;
        int n = 0;
        for (int64_t j = 0; j < M; ++j) {
          if (n == k && !ahead(s[j], j, bs[k - 1], bi[k - 1])) continue;
          int p = n < k ? n++ : k - 1;
          while (p > 0 && ahead(s[j], j, bs[p - 1], bi[p - 1])) { bs[p] = bs[p - 1]; bi[p] = bi[p - 1]; --p; }
          bs[p] = s[j]; bi[p] = j;
        }
        for (; n < k; ++n) { bi[n] = -1; bs[n] = -INFINITY; }
      }
      if (out_rank) {
        const int64_t t = target[qi];
        int64_t r = 1;
        for (int64_t j = 0; j < M; ++j) r += ahead(s[j], j, s[t], t);
        out_rank[qi] = r;
      }
    }
    free(qd); free(s);
  }
  return 0;
}
