// TMA-fed tcgen05 / TMEM similarity scan with fused fusion-weight + top-k / count epilogue
// (north_star kernels 1a + 2).  sm_100a only; hand-written PTX, no CUTLASS.
//
// Orientation: D[128 x N] (fp32, TMEM) = Qblk[128 x D] * Gtile[N x D]^T, both operands bf16,
// K-major, staged in shared memory by TMA with the 128-byte swizzle:
//   A = 128 query rows (rows past the batch are never loaded; their accumulator lanes are ignored)
//   B = N_TILE gallery rows (TMA zero-fills rows past the end of the shard and columns past D)
// so one TMEM lane == one query and the epilogue thread that owns the lane walks the gallery
// columns.  With two galleries (T2I + T2T) each gets its own accumulator and the epilogue forms
// w_a*acc_a + w_b*acc_b in fp32.  Accumulators are double buffered in TMEM (2 x G x N_TILE <= 512
// columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation),
// warps 2..9 = epilogue: warp w reads TMEM lane quadrant w%4 and column half (w-2)/4 of the tile,
// so every scheduler hosts two epilogue warps and a query's candidates come as two lists per CTA.
// One pipeline stage holds the query chunk ONCE plus the matching chunk of every gallery (the
// A operand is shared by the T2I and T2T MMAs).
// Work: U persistent units (CTAs / pairs / quads, <= #SMs) share n_qb query blocks x M gallery rows in ROWS, not
// tiles.  The gallery is cut into floor(U / n_qb) full stripes of L = ceil32(n_qb * M / U) rows; unit u = s * n_qb + qb
// scans stripe s for query block qb, so the n_qb units of a stripe read the SAME rows at the same time (one DRAM read,
// n_qb - 1 L2 hits -- with unsynchronised ranges every block re-read the gallery from DRAM).  The rows left over form
// a remainder stripe that the remaining U mod n_qb units share as one flat (block, row) range.  A unit walks its rows
// in tiles of n_tile; the tile at the end of a range is PARTIAL: its MMA is issued with N = the rows left (rounded up
// to 32) and its rows arrive as 16-row TMA boxes, so every unit does the same tensor work to within 32 rows whatever
// M / n_tile leaves over (C2: 9.08 tile-equivalents on each of 74 pairs instead of 9 or 10 on 72).  A unit emits one
// K-entry candidate list per query and column half for every query block it touches ("part slot"), merged later by
// select.cuh.
//
// Epilogue per element: 1-2 FFMA for the fusion weights and one compare against the thread's running
// threshold.  Survivors are appended to a small per-thread buffer in shared memory; when any lane's
// buffer is half full the whole warp folds its buffers into per-thread sorted top-K lists that live
// in REGISTERS (K <= 32, branch-free compare/select insert), so list maintenance runs in lock-step
// instead of diverging per lane, and the score matrix never exists outside TMEM.
// Shared threshold: a list that starts from -inf takes ~K ln(n / K) inserts for n scores, and a query has dozens of
// lists (two per unit touching its block).  Every FULL list publishes its K-th score per query in global memory
// (red.max on the order-preserving image); every list prunes with the maximum it has seen there (re-read once per
// tile, the load in flight while the warp waits for the accumulator).  A row dropped that way has K better rows in ONE
// list whose final last key is at least the threshold used -- exactly the event "a list rejected rows below its last
// key" that the selection's certificate already accounts for (select.cuh: every full list's last key enters the
// rejection bound), so which rows a list keeps now depends on timing, the certified top-k does not.
// Roofline: tensor pipe for batch >= ~250 (2*B*G*M*D flop), HBM below (G*M*D*2 bytes).
//
// CTA pairs (PAIR = true, batches above 128 queries): two CTAs of one cluster run ONE
// tcgen05.mma.cta_group::2 with M = 256 queries (128 per CTA) and N = 256 gallery columns; each CTA
// stages its own 128 query rows and only HALF of the gallery chunk (CTA 0: T2I rows / first 128 rows,
// CTA 1: T2T rows / second 128 rows), so a pipeline stage is 32 KB instead of 48 KB and the bytes
// pulled through L2 per flop drop by a third -- the single-CTA kernel was bound by L2->SM bandwidth
// (ncu: tensor pipe 61 % of active cycles, 11.5 TB/s of TMA reads).  The leader CTA's elected thread
// issues every MMA; tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both
// CTAs, the peer's TMA loads complete on the leader's full barrier, and both CTAs' epilogue warps
// arrive on the leader's accumulator-empty barrier.
//
// Quads (CL = 4, long scans of >= 1024 queries): two pairs of one cluster take adjacent 256-query blocks and the
// SAME gallery tiles; each CTA fetches a quarter of the gallery chunk with a cta_group::2 multicast TMA load whose
// destinations are the two CTAs holding that half, so the tile crosses L2 once per quad; a stage is free when the
// MMAs of BOTH pairs have retired (empty barriers take two arrivals, commit mask 0xF).
//
// Merged double stage (two galleries, equal weights, pairs / quads): a stage holds the query chunk once plus the
// matching chunk of both galleries; 8 MMAs accumulate q.(a) and q.(b) into the same accumulator.
//
// Issue discipline: the producer and the MMA warp run their loops with all 32 lanes (uniform control flow, so
// addresses and descriptors live in uniform registers) and elect.sync picks the issuing lane around the
// TMA / MMA / commit instructions only.  Issuing from `if (lane == 0)` put a vector->uniform register waterfall
// (ELECT, R2UR.BROADCAST x5, BRA.U.ANY) in front of every UTCHMMA and cost a third of the tensor pipe.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <vector>
#include "scan_warp.cuh"

namespace kemr {

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMmaThreads = 64 + kEpiThreads;
constexpr int kBufCap = 16;            // append-buffer entries per epilogue thread
constexpr int kBufTrigger = 8;         // fold buffers into the lists when any lane holds this many (+8 new fit)
constexpr int kBlockM = 128;           // queries per block == TMEM lanes
constexpr int kBlockK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kSmemBudget = 227 * 1024;
constexpr int kMaxStages = 16;
constexpr int kTraceBase = 8192, kTraceLen = 96;   // KEMR_MMA_DEBUG: per-stage timestamps of CTA 0/1

struct MmaPlan {
  int parts = 0;          // candidate lists per query (2 per CTA touching its block)
  int q_pad = 0;          // queries rounded up to kBlockM
  int n_tile = 0;         // gallery rows per MMA tile (128 with two galleries, else 256)
  int n_qb = 0, n_t = 0;  // query blocks, gallery tiles
  int ctas = 0;
  int stages = 0;
  int kc = 0;             // K chunks of 64
  int a_rows = 0;         // query rows actually loaded per block
  int K = 0;              // list length (8, 16, 24, 32)
  int pair = 0;           // 1: CTA pairs (cta_group::2), 256 queries per block
  int cl = 1;             // CTAs per cluster: 1, 2 (one pair) or 4 (two pairs on adjacent query blocks sharing the gallery tile by TMA multicast)
  int gran = 32;          // unit boundaries are multiples of this many rows inside a query block (n_tile: no partial tiles)
  int s_full = 0;         // full stripes (one unit per stripe and query block)
  int vq = 1;             // virtual parts per query block: lists are flushed and restarted at these boundaries
  int all_slots = 0;      // 1: every part slot of every query row is written (no memset needed)
  int two = 0;            // 1: two accumulators (T2I, T2T) with their own weights
  int merged = 0;         // 1: two galleries, equal weights: one accumulator over 2*kc K chunks
  int ds = 0;             // 1: merged + CTA pairs: query chunk staged once per K step for both galleries
  int q_blk = 0;          // queries per block (128, or 256 for CTA pairs)
  size_t smem = 0;
};

inline bool mma_built() { return true; }
inline bool mma_supported(int D, int K) { return D % 8 == 0 && D >= 8 && D <= kMaxD && K >= 1 && K <= kMaxKSel; }

static thread_local char g_mma_error[256] = "";
inline const char* mma_last_error() { return g_mma_error; }

// First position (block * len + row) of unit u when `rtot` = blocks * len rows of work are cut into `units` ranges:
// the even split, snapped down to a multiple of `gran` rows inside its block.  unit_begin(units) == rtot.
__host__ __device__ inline long long unit_begin(long long rtot, int units, long long len, int gran, int u) {
  const long long x = rtot * u / units;
  const long long qb = x / len, r = x - qb * len;
  return qb * len + r - r % gran;
}
// the unit whose range holds position `pos` (largest u with unit_begin(u) <= pos; ranges may be empty)
__host__ __device__ inline int unit_of(long long rtot, int units, long long len, int gran, long long pos) {
  int lo = 0, hi = units - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (unit_begin(rtot, units, len, gran, mid) <= pos) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// How the units share the work (see the header comment): `s_full` = floor(units / n_qb) stripes, one unit per (stripe,
// block); stripe s starts at row floor_gran(s * n_qb * M / (units - 2 r / 3)), r = units mod n_qb (cumulative rounding:
// stripe lengths differ by at most `gran`); the rows from the end of the last full stripe on are the remainder
// stripe, shared by the r units past s_full * n_qb.  A remainder unit counts as A THIRD of a unit: it walks several
// query blocks (its lists restart cold each time) and nobody reads its rows at the same time (DRAM latency instead of
// L2 hits) -- with equal shares the remainder units of C1 ran 1.67x longer than the others and set the kernel time,
// with half shares still 1.14x.
// Without remainder units the last full stripe ends at M.
struct WorkSplit {
  long long M;        // gallery rows
  int n_qb, units, s_full, gran;
  __host__ __device__ int full_units() const { return s_full * n_qb; }
  __host__ __device__ long long stripe_begin(int s) const {
    if (s >= s_full && units == s_full * n_qb) return M;
    const long long x = (long long)s * n_qb * M * 3 / (3ll * units - 2ll * (units - s_full * n_qb));
    const long long b = x - x % gran;
    return b < M ? b : M;
  }
  __host__ __device__ long long rem_base() const { return stripe_begin(s_full); }
  __host__ __device__ long long rem_len() const { return M - rem_base(); }
};
// Rows per tile of a segment of `rows` rows: the fewest tiles of at most n_tile rows, of (nearly) EQUAL size -- a
// 290-row segment is walked as 160 + 130 rows, not 256 + 34: the epilogue of the first tile then hides behind the
// loads of the second, and the tail after the last byte arrives is one short tile (small batches at 43 k rows).
__host__ __device__ inline int seg_tile_rows(long long rows, int n_tile, int gran) {
  if (gran >= n_tile || rows <= 0) return n_tile;
  const long long nt = (rows + n_tile - 1) / n_tile;
  const long long t = ((rows + nt - 1) / nt + gran - 1) / gran * gran;
  return (int)(t < n_tile ? t : n_tile);
}

// One unit's walk: positions [p_lo, p_hi) of a flat space of blocks of `mod` rows; position p is row base + p % mod of
// query block qb0 + p / mod.
struct UnitWalk { long long base, mod, p_lo, p_hi; int qb0; };

__host__ __device__ inline UnitWalk unit_walk(const WorkSplit& w, int u) {
  UnitWalk k;
  if (u < w.full_units()) {
    const int s = u / w.n_qb;
    k.base = w.stripe_begin(s);
    const long long len = w.stripe_begin(s + 1) - k.base;
    k.mod = len > 0 ? len : 1; k.p_lo = 0; k.p_hi = len > 0 ? len : 0; k.qb0 = u % w.n_qb;
  } else {
    const int nr = w.units - w.full_units(), j = u - w.full_units();
    const long long len = w.rem_len();
    k.base = w.rem_base(); k.mod = len > 0 ? len : 1; k.qb0 = 0;
    const long long rtot = (long long)w.n_qb * len;
    k.p_lo = len > 0 ? unit_begin(rtot, nr, len, w.gran, j) : 0;
    k.p_hi = len > 0 ? unit_begin(rtot, nr, len, w.gran, j + 1) : 0;
  }
  return k;
}
// ordinal of unit u among the units that touch query block qb, in row order (full stripes first, then the remainder units)
__host__ __device__ inline int unit_ordinal(const WorkSplit& w, int u, int qb) {
  if (u < w.full_units()) return u / w.n_qb;
  const int nr = w.units - w.full_units();
  const long long len = w.rem_len();
  const int first = unit_of((long long)w.n_qb * len, nr, len, w.gran, (long long)qb * len);
  return w.s_full + (u - w.full_units() - first);
}
// number of units that touch query block qb
__host__ __device__ inline int units_of_block(const WorkSplit& w, int qb) {
  int n = w.s_full;
  const int nr = w.units - w.full_units();
  const long long len = w.rem_len();
  if (nr > 0 && len > 0) {
    const long long rtot = (long long)w.n_qb * len;
    n += unit_of(rtot, nr, len, w.gran, (long long)(qb + 1) * len - 1) - unit_of(rtot, nr, len, w.gran, (long long)qb * len) + 1;
  }
  return n;
}
inline WorkSplit make_split(long long M, int n_qb, int units, int gran) {
  WorkSplit w;
  w.M = M; w.n_qb = n_qb; w.units = units; w.gran = gran;
  w.s_full = units / n_qb;
  return w;
}

constexpr int kPlanMaxParts = 304;     // candidate lists per query any plan may use (workspace bound: 2*148 + 8)

inline int mma_make_plan(int Q, int64_t M, int D, int G, int K, int mode, int sms, int quads, bool equal_weights, MmaPlan* p) {
  (void)mode;
  p->kc = (D + kBlockK - 1) / kBlockK;
  // Two galleries with EQUAL fusion weights: w*(q.a) + w*(q.b) = w*(q.a + q.b), so both galleries'
  // K chunks accumulate into one accumulator of 256 gallery rows and the epilogue sees half as many
  // scores per flop.  Different weights keep one accumulator per gallery (128 rows each).
  static const bool no_merge = getenv("KEMR_MMA_NO_MERGE") != nullptr;
  p->merged = (G == 2 && equal_weights && !no_merge) ? 1 : 0;
  p->two = (G == 2 && !p->merged) ? 1 : 0;
  // Measured and rejected on B200: A operand from TMEM (tcgen05.mma TS form) costs ~70 cycles + N/2
  // per MMA whatever N is (TMEM read of the 128x16 A tile), 5x slower at N = 32 -- removed.
  static const char* force_pair = getenv("KEMR_MMA_PAIR");          // experiments: 0 = never, 1 = always
  p->pair = force_pair ? (force_pair[0] == '1') : (Q > kBlockM);
  if (sms < 2) p->pair = 0;
  // Clusters of FOUR CTAs (two pairs): the pairs take adjacent 256-query blocks and the same gallery tiles, and
  // every CTA fetches only a quarter of the tile chunk, multicast to the CTA of the other pair that needs the same
  // half -> 24 KB instead of 32 KB through L2 per CTA and stage (the pair kernel reaches the L2 slice limit of
  // ~1 sector/clk/slice, ncu).  `quads` = clusters of four the device can hold at once (33 on a B200: GPC shapes
  // leave 16 SMs unusable), so this pays only for long L2-bound scans: measured +7 % on 4096 q x 1.25 M rows,
  // -7 % on 1000 q x 43 k rows x 2 galleries (fewer SMs, 11 instead of 10 tiles per unit).
  const int n_qb256 = p->pair ? (Q + 2 * kBlockM - 1) / (2 * kBlockM) : 0;
  const int64_t nt256 = (M + (p->two ? 128 : 256) - 1) / (p->two ? 128 : 256);
  static const char* force_cl = getenv("KEMR_MMA_CL");              // experiments: 2 = pairs only, 4 = always quads
  bool quad = p->pair && quads >= 1 && n_qb256 >= 4 && n_qb256 % 2 == 0 && nt256 >= 1024;
  if (force_cl && p->pair && quads >= 1 && n_qb256 >= 2) quad = force_cl[0] == '4';
  p->cl = quad ? 4 : (p->pair ? 2 : 1);
  p->q_blk = quad ? 4 * kBlockM : (p->pair ? 2 * kBlockM : kBlockM);     // queries per scheduling block (per cluster)
  p->n_tile = p->two ? 128 : 256;
  p->n_qb = (Q + p->q_blk - 1) / p->q_blk;
  p->q_pad = p->n_qb * p->q_blk;
  const int64_t nt = (M + p->n_tile - 1) / p->n_tile;
  if (nt > (1ll << 30)) return 1;
  p->n_t = (int)nt;
  const int units = quad ? quads : (p->pair ? sms / 2 : sms);       // persistent CTAs, pairs or quads
  // Partial tiles (MMA with a smaller N, 16-row TMA boxes) exist for the single-accumulator kernels of one CTA or one
  // pair; quads and the two-accumulator kernels keep tile-aligned boundaries.
  static const bool no_partial = getenv("KEMR_MMA_NO_PARTIAL") != nullptr;     // experiments: tile-aligned boundaries everywhere
  p->gran = (quad || p->two || no_partial) ? p->n_tile : (p->pair ? 32 : 16);      // MMA N: multiples of 16 (one CTA) / 32 (pair)
  const int64_t rtot = (int64_t)p->n_qb * M;
  const int nu = (int)std::min<int64_t>(units, std::max<int64_t>(1, rtot / p->n_tile));   // at least a tile's worth of rows per unit
  const WorkSplit ws = make_split(M, p->n_qb, nu, p->gran);
  p->s_full = ws.s_full;
  int parts = 1;
  bool same = true;
  for (int qb = 0; qb < p->n_qb; ++qb) {
    const int n = units_of_block(ws, qb);
    if (qb && n != parts) same = false;
    parts = std::max(parts, n);
  }
  for (int u = 0; u < nu; ++u) {
    const UnitWalk k = unit_walk(ws, u);
    if (k.p_lo >= k.p_hi) same = false;                     // an empty unit leaves a hole in the part slots
  }
  p->ctas = p->cl * nu;
  p->a_rows = (p->pair || Q >= kBlockM) ? kBlockM : (Q + 7) / 8 * 8;
  // Candidate lists.  A query's rows are cut into segments (unit boundaries, plus `vq` virtual
  // boundaries per block when more are needed); every segment yields two K-entry lists (one per
  // column half).  Each list keeps the exact top-K of its rows, so the K_sel best rows of the query
  // survive unless K or more of them fall into one list.  With the winners spread at random over L
  // lists that happens with probability <= L * P[Poisson(K_sel / L) >= K] per query; (K, L) is chosen
  // so that this stays below 2e-6 (a list that does overflow is caught by the select kernel's
  // certificate and the query is re-run), taking the cheapest of K = 8 and 16 (cost ~ keys L*K times
  // the insert cost ~K), and K = 32 only when neither fits.  Large k (top-100) therefore runs with short register lists
  // and MANY parts instead of long lists.
  const int span = parts;                          // natural segments per query block
  auto overflow_prob = [](int ksel, int L, int Kc) {
    if (ksel < Kc) return 0.0;
    const double m = (double)ksel / L;
    double term = exp(-m);
    for (int i = 1; i <= Kc; ++i) term *= m / i;
    double sum = term;
    for (int i = Kc + 1; i < Kc + 64; ++i) { term *= m / i; sum += term; }
    return sum * L;
  };
  const int seg_cap = (int)std::min<int64_t>(std::max<int64_t>(span, (int64_t)p->n_t), kPlanMaxParts / 2 - span + 1);
  int Ksel = 0, vq = 1;
  long long best_keys = 1ll << 60;
  static const char* force_k = getenv("KEMR_MMA_KLIST");            // experiments: list length 8 / 16 / 32
  for (int Kc = 8; Kc <= 32; Kc *= 2) {
    if (force_k && atoi(force_k) != Kc) continue;
    if (Kc == 32 && Ksel) break;
    for (int seg = span; seg <= std::max(span, seg_cap); ++seg) {
      const int segs = seg == span ? span : seg + span - 1;            // list slots when virtual parts are added
      if (2 * segs > kPlanMaxParts) break;
      if (overflow_prob(K, 2 * seg, Kc) <= 2e-6) {
        // cost ~ keys the selection reads (L * K) times the insert cost (~K).  Virtual parts restart the lists, but not
        // cold: the per-query shared threshold (thr_pub) hands a restarted list the best K-th score any list of the
        // query has reached.  C1 (4300 q x 43 k x 512): K = 16 with 12 natural lists 199 us; K = 8 with 40 lists 181 us
        // with the shared threshold, 211 us without it (profiles/r02_session_m_stdout.txt).
        const long long keys = 2ll * segs * Kc * Kc;
        if (keys < best_keys || (keys == best_keys && seg == span)) { best_keys = keys; Ksel = Kc; vq = seg == span ? 1 : seg; }
        break;
      }
      if (seg >= p->n_t) break;                    // cannot cut finer than one tile per segment
    }
  }
  if (!Ksel) return 1;
  p->K = Ksel;
  p->vq = vq;
  p->parts = 2 * (vq + span - 1);
  p->all_slots = (vq == 1 && same) ? 1 : 0;     // every block is covered by the same number of non-empty units
  static const bool no_ds = getenv("KEMR_MMA_NO_DS") != nullptr;
  p->ds = (p->merged && p->pair && !no_ds) ? 1 : 0;
  const size_t stage = (size_t)kBlockM * 128 + (p->pair ? (size_t)128 * 128 * (p->ds ? 2 : 1) : (size_t)256 * 128);
  const size_t epi = (size_t)kBufCap * kEpiThreads * 8;
  p->stages = (int)std::min<size_t>(kMaxStages, (kSmemBudget - 2048 - epi) / stage);
  p->smem = (size_t)p->stages * stage + epi + 1024;
  return p->stages >= 2 ? 0 : 1;
}

// ------------------------------------------------------------------------------------ PTX wrappers
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// one lane of the (converged) warp; always the same lane for a full mask
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA pair (cta_group::2) forms
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory object in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope): a
// cluster-scope release made the issuing thread wait ~1.5-2.5k cycles for its outstanding TMA loads.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA's share of a pair's stage; completion bytes go to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// the same, multicast: the box lands at the same offset in every CTA of `mask`, and each destination's bytes are
// signalled on the barrier at `leader_bar`'s offset in that destination's pair leader (CUTLASS SM100_TMA_2SM_LOAD_MULTICAST)
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this shared-memory offset in every CTA of `mask` when the MMAs retire
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct MmaArgs {
  ScanArgs s;
  int n_tile;       // gallery rows per tile (128 with two accumulators, else 256)
  int merged;       // both galleries accumulate into ONE accumulator (equal weights): 2*kc K chunks
  int ds;           // merged, CTA pairs: a stage holds the query chunk ONCE plus the matching chunk of BOTH galleries (8 MMAs)
  int n_qb, n_t, stages, kc, kc_total, a_rows, parts, q_pad, q_blk, gran, vq;
  WorkSplit split;  // how the units share the (query block, gallery row) work
  long long* dbg;   // optional [ctas][16] cycle counters + stage trace (debug build, KEMR_MMA_DEBUG=1)
  int epi_variant;  // short lists: 2 = per-lane predicated appends on RAW accumulators (one accumulator, positive weight), 1 = on weighted scores; 0 = max tree + vote (long scans); KEMR_MMA_EPI overrides.  (Measured and rejected: a survivor bit mask per 16 columns + select-tree extraction -- C2 107 -> 121 us, C1 318 -> 393 us in the same build.)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

// bounded wait that optionally accumulates the cycles spent waiting (debug counters)
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, bool dbg, long long& acc) {
  if (!dbg) { ptx::mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  ptx::mbar_wait(bar, parity);
  acc += clock64() - t0;
}

// sorted (descending) per-thread candidate list in registers; rows arrive in increasing order, so
// strict '>' keeps the lower index ahead among equal scores
template <int K>
struct RegList {
  float sc[K];
  uint32_t ix[K];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int p = 0; p < K; ++p) { sc[p] = -INFINITY; ix[p] = 0xffffffffu; }
  }
  __device__ __forceinline__ float threshold() const { return sc[K - 1]; }
  __device__ __forceinline__ void insert(float s, uint32_t r) {
    bool up_next = true;                       // s > sc[K-1] is the caller's precondition
#pragma unroll
    for (int p = K - 1; p >= 1; --p) {
      const bool up = s > sc[p - 1];           // new entry passes position p-1 -> p-1 moves down to p
      const float nsc = up ? sc[p - 1] : (up_next ? s : sc[p]);
      const uint32_t nix = up ? ix[p - 1] : (up_next ? r : ix[p]);
      sc[p] = nsc; ix[p] = nix;
      up_next = up;
    }
    if (up_next) { sc[0] = s; ix[0] = r; }
  }
};

__device__ __forceinline__ float max8(const float* v) {
  return fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, float s, uint32_t r) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(s)), "r"(r) : "memory");
}
__device__ __forceinline__ void ld_shared_v2(uint32_t addr, float& s, uint32_t& r) {
  uint32_t u;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u), "=r"(r) : "r"(addr) : "memory");
  s = __uint_as_float(u);
}

// Smallest of the K group maxima of one lane's HALF tile of fused scores (HC columns from `acc`, cut into K groups of
// HC / K columns): at least K of these scores reach it.
template <int K, bool TWO, int HC>
__device__ __forceinline__ float warm_floor(uint32_t acc, float w0, float w1) {
  constexpr int g = (HC % K == 0 && HC / K >= 1 && HC / K <= 16) ? HC / K : 1;     // columns per group
  float v = INFINITY;
  for (int c0 = 0; c0 < HC; c0 += 16) {
    uint32_t xa[16], xb[16];
    tmem_ld16(acc + (uint32_t)c0, xa);
    if (TWO) tmem_ld16(acc + 128u + (uint32_t)c0, xb);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += g) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < g; ++j) {
        float s = w0 * __uint_as_float(xa[j0 + j]);
        if (TWO) s = fmaf(w1, __uint_as_float(xb[j0 + j]), s);
        m = fmaxf(m, s);
      }
      v = fminf(v, m);
    }
  }
  return v;
}

// K = register list length; PAIR = CTA pairs (cta_group::2, 256 queries per block); TWO = two
// accumulators (T2I, T2T) of 128 columns each, fused in the epilogue with their own weights
// (otherwise ONE accumulator of 256 gallery rows: single gallery, or both galleries with equal
// weights accumulated over 2*kc K chunks).
// MODE (kModeTopk / kModeCount / kModeDense) is a template parameter: one epilogue per kernel keeps the loop body of
// the eight epilogue warps inside the instruction cache (with every epilogue inlined C2 lost 11 %, C1 25 %).
template <int K, int CL, bool TWO, int MODE>
__global__ void __launch_bounds__(kMmaThreads, 1)
scan_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g0,
                const __grid_constant__ CUtensorMap map_g1, const __grid_constant__ CUtensorMap map_s0,
                const __grid_constant__ CUtensorMap map_s1, MmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_mma_raw[];
  // identical shared-memory layout in both CTAs of a pair (the MMA addresses the peer by offset)
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_mma_raw + 1023) & ~(uintptr_t)1023);
  constexpr bool PAIR = CL >= 2;                                // CTA pairs (cta_group::2)
  constexpr bool QUAD = CL == 4;                                // two pairs per cluster sharing the gallery tile
  constexpr int n_tile = TWO ? 128 : 256;                       // gallery rows per tile
  constexpr uint32_t a_bytes = (uint32_t)kBlockM * 128u;
  constexpr uint32_t b_box = PAIR ? 128u : (uint32_t)n_tile;     // gallery rows per CTA and stage
  constexpr uint32_t b_bytes = b_box * 128u;
  const bool ds = PAIR && !TWO && a.ds != 0;                    // merged double stage: A + B(T2I) + B(T2T)
  const uint32_t stage_bytes = a_bytes + (PAIR ? (ds ? 2u * b_bytes : b_bytes) : 256u * 128u);
  const uint32_t buf_u32 = ptx::smem_u32(smem + (size_t)a.stages * stage_bytes);      // [kBufCap][kEpiThreads] x 8 B
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes + (size_t)kBufCap * kEpiThreads * 8);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;    // rank in the cluster
  const uint32_t rank = crank & 1u;                             // rank in the pair; 0 = leader (issues the MMAs)
  const uint32_t pc = crank >> 1;                               // which pair of a quad (its 256-query block)
  const uint32_t lead = crank & ~1u;                            // cluster rank of this pair's leader
  const int unit = (int)(blockIdx.x / CL);                      // persistent CTA, pair or quad
#ifdef KEMR_DEBUG
  const bool dbg = a.dbg != nullptr;
  if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.dbg[blockIdx.x * 16 + 8] = (long long)t; }
#else
  constexpr bool dbg = false;          // role counters and the stage trace exist in the debug build only
#endif
  // this unit's share of the work: a stripe of one query block, or a flat range of the remainder stripe
  const UnitWalk uw = unit_walk(a.split, unit);
  const long long p_lo = uw.p_lo, p_hi = uw.p_hi;
  // partial tiles: the single-accumulator kernels of one CTA / one pair, when the plan cut the ranges finer than tiles
  const bool partial_ok = !TWO && !QUAD && a.gran < n_tile;
  // Walk of a range, identical in the three roles: segments (one per query block touched) of tiles; the last tile of
  // a segment may be partial.  KEMR_FOR_TILES(body) runs `body` with qb, row0 (first row of the tile inside the
  // block) and ncols (rows of the tile that belong to this unit) in scope.
#define KEMR_FOR_TILES(...)                                                                   \
  for (long long p__ = p_lo; p__ < p_hi;) {                                                   \
    const long long b__ = p__ / uw.mod;                                                       \
    const int qb = uw.qb0 + (int)b__;                                                         \
    const long long blk0__ = b__ * uw.mod;                                                    \
    const long long rend__ = (p_hi < blk0__ + uw.mod ? p_hi : blk0__ + uw.mod) - blk0__;      \
    const int tsz__ = seg_tile_rows(rend__ - (p__ - blk0__), n_tile, partial_ok ? a.gran : n_tile); \
    for (long long r__ = p__ - blk0__; r__ < rend__; r__ += tsz__) {                          \
      const long long row0 = uw.base + r__;                                                   \
      const int ncols = (int)(rend__ - r__ < (long long)tsz__ ? rend__ - r__ : (long long)tsz__); \
      __VA_ARGS__                                                                             \
    }                                                                                         \
    p__ = blk0__ + rend__;                                                                    \
  }

  if (threadIdx.x == 0) {
    // full: one arrival (the leader's expect_tx covers both CTAs' bytes; the peer's TMA loads complete
    // on the leader's barrier); tempty: epilogue warps of both CTAs
    // empty: one arrival per pair of the cluster (a quad's CTAs write into each other's stages)
    for (int i = 0; i < a.stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], QUAD ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], PAIR ? 2 * kEpiWarps : kEpiWarps); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&map_q); ptx::prefetch_tmap(&map_g0);
    if (a.s.G > 1) ptx::prefetch_tmap(&map_g1);
  }
  if (warp == 1) { if (PAIR) ptx::tmem_alloc_pair(tmem_ptr, 512); else ptx::tmem_alloc(tmem_ptr, 512); }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // programmatic dependent launch: the selection kernel queued behind this scan may take its SM slots as they free up
  // (its CTAs block in griddepcontrol.wait until this whole grid has finished and its lists are visible)
  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ================================================================= TMA producer
    // (whole warp in the loop, one elected lane issues -- see the MMA issuer below)
    {
      int stage = 0; uint32_t phase = 0;
      long long w_empty = 0; const long long t_begin = dbg ? clock64() : 0;
      int tr = 0;
      KEMR_FOR_TILES({
        // rows of the gallery chunk this CTA stages: a full tile, or (single accumulator) the rows left, to 32
        const int nmma = partial_ok ? min(256, (ncols + (PAIR ? 31 : 15)) & (PAIR ? ~31 : ~15)) : 256;
        const int nb = PAIR ? (TWO ? 128 : nmma >> 1) : (TWO ? 128 : nmma);
        const bool whole = QUAD || nb == (int)b_box;
        const uint32_t g_bytes = (uint32_t)nb * 128u * ((TWO && !PAIR) || ds ? 2u : 1u);     // gallery bytes per CTA and stage
        const uint32_t tx = PAIR ? 2u * (a_bytes + g_bytes) : (uint32_t)a.a_rows * 128u + g_bytes;
        const int brow = (int)row0 + ((PAIR && !TWO) ? (int)rank * nb : 0);
        for (int kcc = 0; kcc < a.kc_total; ++kcc) {
          const int g = kcc >= a.kc ? 1 : 0;              // merged mode: second gallery's chunks follow the first's
          const int kx = (kcc - g * a.kc) * kBlockK;
          mbar_wait_dbg(&empty_bar[stage], phase ^ 1, dbg, w_empty);
          if (dbg && blockIdx.x < 2 && lane == 0 && tr < kTraceLen) a.dbg[kTraceBase + (blockIdx.x * 4 + 0) * kTraceLen + tr] = clock64();
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          if (ptx::elect_one()) {
            if (PAIR) {
              // this CTA's 128 query rows + its half of the gallery chunk; bytes land on the leader's barrier
              const uint32_t lbar = ptx::map_to_cta(ptx::smem_u32(&full_bar[stage]), lead);
              if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], tx);
              ptx::tma_load_2d_pair(sa, &map_q, lbar, kx, qb * a.q_blk + (int)crank * kBlockM);
              for (int gi = 0; gi < (ds ? 2 : 1); ++gi) {
                const bool second = TWO ? rank != 0 : (ds ? gi : g) != 0;
                unsigned char* sb = sa + a_bytes + (uint32_t)gi * b_bytes;
                if (QUAD) {
                  // this CTA's quarter of the chunk (64 rows), multicast to the CTA of the other pair holding the same half
                  ptx::tma_load_2d_pair_mc(sb + pc * (64u * 128u), second ? &map_g1 : &map_g0, lbar, kx, brow + (int)pc * 64,
                                           (uint16_t)((1u << rank) | (4u << rank)));
                } else if (whole) {
                  ptx::tma_load_2d_pair(sb, second ? &map_g1 : &map_g0, lbar, kx, brow);
                } else {
                  for (int r16 = 0; r16 < nb; r16 += 16)      // partial tile: 16-row boxes, same swizzled layout
                    ptx::tma_load_2d_pair(sb + (uint32_t)r16 * 128u, second ? &map_s1 : &map_s0, lbar, kx, brow + r16);
                }
              }
            } else {
              ptx::mbar_expect_tx(&full_bar[stage], tx);
              ptx::tma_load_2d(sa, &map_q, &full_bar[stage], kx, qb * kBlockM);
              if (TWO) {
                ptx::tma_load_2d(sa + a_bytes, &map_g0, &full_bar[stage], kx, brow);
                ptx::tma_load_2d(sa + a_bytes + b_bytes, &map_g1, &full_bar[stage], kx, brow);
              } else if (whole) {
                ptx::tma_load_2d(sa + a_bytes, g ? &map_g1 : &map_g0, &full_bar[stage], kx, brow);
              } else {
                for (int r16 = 0; r16 < nb; r16 += 16)
                  ptx::tma_load_2d(sa + a_bytes + (uint32_t)r16 * 128u, g ? &map_s1 : &map_s0, &full_bar[stage], kx, brow + r16);
              }
            }
          }
          __syncwarp();
          if (dbg && blockIdx.x < 2 && lane == 0 && tr < kTraceLen) a.dbg[kTraceBase + (blockIdx.x * 4 + 1) * kTraceLen + tr] = clock64();
          if (dbg && blockIdx.x < 2 && tr < kTraceLen) ++tr;
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      })
      if (dbg && lane == 0) { a.dbg[blockIdx.x * 16 + 0] = w_empty; a.dbg[blockIdx.x * 16 + 1] = clock64() - t_begin; }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (leader CTA only)
    // The WHOLE warp walks the loop (warp-uniform control flow and addresses, so descriptors live in uniform
    // registers); one elected lane issues the MMAs and commits.  Issued from a single divergent lane, every MMA
    // sat behind a ~60-cycle register->uniform-register waterfall and the tensor pipe idled between MMAs
    // (ncu: 64 % tensor-pipe activity with all loads and the epilogue switched off).
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0;
      // The gallery chunk(s) of a stage form ONE K-major tile of 256 rows (T2I rows then T2T rows, or
      // 256 rows of one gallery; split across the two CTAs in pair mode): a single MMA with N = 256
      // fills the whole accumulator buffer and reads the query chunk once.
      const uint64_t adesc0 = umma_desc_sw128(ptx::smem_u32(smem));
      const uint64_t bdesc0 = umma_desc_sw128(ptx::smem_u32(smem) + a_bytes);
      const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);          // descriptor start-address units (16 B)
      long long it = 0;
      long long w_full = 0, w_tempty = 0; const long long t_begin = dbg ? clock64() : 0;
      int tr = 0;
      KEMR_FOR_TILES({
        const int buf = (int)(it & 1);
        const uint32_t bphase = (uint32_t)((it >> 1) & 1);
        // a partial tile (end of this unit's range) fills only the accumulator columns of its rows
        const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, partial_ok ? min(256, (ncols + (PAIR ? 31 : 15)) & (PAIR ? ~31 : ~15)) : 256);
        mbar_wait_dbg(&tempty_bar[buf], bphase ^ 1, dbg, w_tempty);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
        for (int kcc = 0; kcc < a.kc_total; ++kcc) {
          mbar_wait_dbg(&full_bar[stage], phase, dbg, w_full);
          if (dbg && blockIdx.x == 0 && lane == 0 && tr < kTraceLen) a.dbg[kTraceBase + 2 * kTraceLen + tr] = clock64();
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)stage * stage_step;
          const uint64_t bdesc = bdesc0 + (uint64_t)stage * stage_step;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (PAIR) ptx::mma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kcc | k) ? 1u : 0u);
              else ptx::mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kcc | k) ? 1u : 0u);
            }
            if (PAIR && ds) {
              // same query chunk against the second gallery's chunk, same accumulator
              const uint64_t bdesc1 = bdesc + (uint64_t)(b_bytes >> 4);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                ptx::mma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc1 + (uint64_t)(k * 2), idesc, 1u);
            }
            // frees the smem stage (in every CTA of the cluster) when these MMAs retire
            if (PAIR) ptx::mma_commit_pair(&empty_bar[stage], QUAD ? (uint16_t)0xF : (uint16_t)0x3); else ptx::mma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (dbg && blockIdx.x == 0 && lane == 0 && tr < kTraceLen) { a.dbg[kTraceBase + 3 * kTraceLen + tr] = clock64(); }
          if (dbg && blockIdx.x == 0 && tr < kTraceLen) ++tr;
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        // accumulators of this tile are complete
        if (ptx::elect_one()) {
          if (PAIR) ptx::mma_commit_pair(&tfull_bar[buf], (uint16_t)(3u << lead)); else ptx::mma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        ++it;
      })
      if (dbg && lane == 0) { a.dbg[blockIdx.x * 16 + 2] = w_full; a.dbg[blockIdx.x * 16 + 3] = w_tempty; a.dbg[blockIdx.x * 16 + 4] = clock64() - t_begin; }
    }
  } else {
    // ================================================================= epilogue (warps 2..9)
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;                    // which half of the tile's score columns
    const int et = (warp - 2) * 32 + lane;               // epilogue thread id, 0..255
    const int qrow = (int)crank * kBlockM + quad * 32 + lane;    // query row inside the (cluster's) block
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t my_buf = buf_u32 + (uint32_t)et * 8u;  // append-buffer entry i of this thread: + i * kEpiThreads * 8
    float w0 = a.s.w[0], w1 = a.s.w[1];
    constexpr int mode = MODE;
    constexpr int kHalfCols = n_tile / 2;                // score columns per warp and tile
    RegList<K> list;
    list.reset();
    float thr = INFINITY;
    float tshare = -INFINITY;                            // largest K-th score published by any full list of this query
    // (the kernels that sit at the register limit -- K = 32, K = 16 with two accumulators -- keep the plain epilogue)
    constexpr bool kLean = K > 16 || (TWO && K >= 16);
    unsigned int* const thr_pub = (mode == kModeTopk && !kLean) ? a.s.thr_pub : nullptr;
    float thr_raw = INFINITY;                            // thr / w0 rounded down: raw accumulator <= thr_raw  =>  w0*acc <= thr
    int bcnt = 0;                                        // entries in this thread's append buffer
    int32_t cnt = 0;
    float blo = 0.f, bhi = 0.f;
    int cur_qb = -1, cur_part = -1, c_first = 0, slot = 0;
    bool cold = false;                                   // the part has just begun: its list is empty
    bool qvalid = false;
    bool warp_active = false;                            // some lane of this warp holds a real query of the current block
    int qg = 0;
    long long it = 0;
    long long w_tfull = 0, t_fold = 0; const long long t_begin = dbg ? clock64() : 0;
    const uint32_t tempty0 = PAIR ? ptx::map_to_cta(ptx::smem_u32(&tempty_bar[0]), lead) : 0u;
    const uint32_t tempty1 = PAIR ? ptx::map_to_cta(ptx::smem_u32(&tempty_bar[1]), lead) : 0u;

    // fold every lane's append buffer into its register list, in lock-step
    auto fold = [&]() {
      const long long tf0 = dbg ? clock64() : 0;
      const int nmax = __reduce_max_sync(0xffffffffu, bcnt);
      // the next entry's shared-memory load is in flight while the current one is inserted (ONE insert body: four
      // entries per round, unrolled, cost C1 20 % -- the epilogue loop has to stay inside the instruction cache)
      if (kLean) {
        for (int i = 0; i < nmax; ++i) {
          if (i < bcnt) {
            float s; uint32_t r;
            ld_shared_v2(my_buf + (uint32_t)i * (kEpiThreads * 8u), s, r);
            if (s > thr) { list.insert(s, r); thr = list.threshold(); }
          }
        }
      } else {
        float s_n = -INFINITY; uint32_t r_n = 0;
        if (nmax > 0) ld_shared_v2(my_buf, s_n, r_n);
        for (int i = 0; i < nmax; ++i) {
          const float s = s_n; const uint32_t r = r_n;
          if (i + 1 < nmax) ld_shared_v2(my_buf + (uint32_t)(i + 1) * (kEpiThreads * 8u), s_n, r_n);
          if (i < bcnt && s > thr) {
            list.insert(s, r);
            thr = fmaxf(list.threshold(), tshare);
          }
        }
      }
      thr_raw = __fdiv_rd(thr, w0);
      bcnt = 0;
      // a full list offers its K-th score to the other lists of the query (fire and forget)
      if (thr_pub && qvalid && list.threshold() > -INFINITY) atomicMax(thr_pub + qg, order_f32(list.threshold()));
      if (dbg) t_fold += clock64() - tf0;
    };
    // (Measured and rejected, A/B on one box, profiles/r02_session_ab_epilogue_microopts.txt: the append position as a
    // running shared-memory address instead of counter + multiply-add, and no per-score column bound (rows past the
    // gallery dropped at fold time instead) -- two instructions fewer per score, C2 unchanged, C1 181.7 -> 192.2 us.)
    // predicated, unrolled: every survivor of an 8-column run goes to the append buffer
    auto append8 = [&](const float* v, uint32_t row) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (v[j] > thr) {
          st_shared_v2(my_buf + (uint32_t)bcnt * (kEpiThreads * 8u), v[j], row + (uint32_t)j);
          ++bcnt;
        }
      }
    };
    // single accumulator, positive weight: compare the RAW accumulator with thr / w0 and weight the survivors only
    auto append8_raw = [&](const uint32_t* v, uint32_t row, int nvalid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = __uint_as_float(v[j]);
        if (x > thr_raw && j < nvalid) {
          st_shared_v2(my_buf + (uint32_t)bcnt * (kEpiThreads * 8u), w0 * x, row + (uint32_t)j);
          ++bcnt;
        }
      }
    };
    // write the lists / count of the part that just ended into its slot
    auto flush = [&]() {
      if (cur_part < 0) return;
      if (mode == kModeTopk) {
        fold();
        uint64_t* dst = a.s.part_keys + ((size_t)slot * a.q_pad + qg) * a.s.K;
#pragma unroll
        for (int p = 0; p < K; ++p)
          if (p < a.s.K) dst[p] = list.ix[p] == 0xffffffffu ? 0ull : make_key(list.sc[p], list.ix[p]);
      } else if (mode == kModeCount) {
        a.s.part_count[(size_t)slot * a.q_pad + qg] = cnt;
      }
    };

    KEMR_FOR_TILES({
      if (qb != cur_qb) {
        cur_qb = qb;
        c_first = unit_ordinal(a.split, unit, qb);     // this unit's ordinal among the units of the block, in row order
      }
      // segment ordinal inside the query block: virtual parts + unit boundaries passed so far
      const int ord = (a.vq > 1 ? (int)((row0 * a.vq) / a.s.M) : 0) + c_first;
      const int part = qb * 4096 + ord;
      if (part != cur_part) {
        flush();
        cur_part = part;
        slot = ord * 2 + half;
        qg = qb * a.q_blk + qrow;
        qvalid = qg < a.s.Q;
        warp_active = __any_sync(0xffffffffu, qvalid);   // small batches: the warps of empty lane quadrants only hand the buffer back
        list.reset();
        cold = true;
        tshare = -INFINITY;
        thr = qvalid ? -INFINITY : INFINITY;             // padded query rows never collect candidates
        thr_raw = thr;
        if (a.s.wq[0] && qvalid) { w0 = a.s.wq[0][qg]; w1 = a.s.wq[1][qg]; }   // per-query gate (two accumulators)
        cnt = 0;
        if (mode == kModeCount) { blo = a.s.band_lo[qg]; bhi = a.s.band_hi[qg]; }   // padded rows hold +huge: never count
      }
      const int buf = (int)(it & 1);
      const uint32_t bphase = (uint32_t)((it >> 1) & 1);
      // the query's shared threshold, fetched while the warp waits for the accumulator
      unsigned int tp = 0;
      if (thr_pub && qvalid) tp = __ldcg(thr_pub + qg);
      mbar_wait_dbg(&tfull_bar[buf], bphase, dbg, w_tfull);
      ptx::tc_fence_after();
      if (tp) {
        const float t = unorder_f32(tp);
        if (t > tshare) {
          tshare = t;
          if (t > thr) { thr = t; thr_raw = __fdiv_rd(thr, w0); }
        }
      }
      const bool full_tile = ncols == n_tile;            // warp-uniform
      const uint32_t acc0 = lane_addr + (uint32_t)buf * 256u;
      const int cbeg = half * kHalfCols;
      const int nch = max(0, min(kHalfCols, ncols - cbeg) + 15) >> 4;       // 16-column chunks for this warp

      auto load = [&](int c0, uint32_t (&ra)[16], uint32_t (&rb)[16]) {
        tmem_ld16(acc0 + (uint32_t)c0, ra);
        if (TWO) tmem_ld16(acc0 + 128u + (uint32_t)c0, rb);
      };
      auto process = [&](int c0, const uint32_t (&ra)[16], const uint32_t (&rb)[16]) {
        if (!TWO && mode == kModeTopk && a.epi_variant == 2 && w0 > 0.f) {
          const uint32_t r = (uint32_t)(row0 + c0);
          const int nv = full_tile ? 16 : ncols - c0;            // columns of this chunk inside the gallery
          append8_raw(ra, r, nv);
          if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
          append8_raw(ra + 8, r + 8u, nv - 8);
          if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
          return;
        }
        float sv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float s = w0 * __uint_as_float(ra[j]);
          if (TWO) s = fmaf(w1, __uint_as_float(rb[j]), s);
          sv[j] = s;
        }
        if (mode == kModeTopk) {
          if (!full_tile) {
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = (c0 + j < ncols) ? sv[j] : -INFINITY;
          }
          if (a.epi_variant) {
            // per-lane predicated appends, one fold check per 8 columns (no max trees, no "any survivor" vote)
            const uint32_t r = (uint32_t)(row0 + c0);
            append8(sv, r);
            if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
            append8(sv + 8, r + 8u);
            if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
          } else {
          const float m0 = max8(sv), m1 = max8(sv + 8);
          if (__any_sync(0xffffffffu, fmaxf(m0, m1) > thr)) {
            const uint32_t r = (uint32_t)(row0 + c0);
            if (m0 > thr) append8(sv, r);
            if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
            if (m1 > thr) append8(sv + 8, r + 8u);
            if (__any_sync(0xffffffffu, bcnt >= kBufTrigger)) fold();
          }
          }
        } else if (mode == kModeCount) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s = sv[j];
            const bool in = full_tile || c0 + j < ncols;
            if (in && s > bhi) ++cnt;
            else if (in && s >= blo) {
              const unsigned int slot = atomicAdd(a.s.amb_counter, 1u);
              if (slot < a.s.amb_cap) { a.s.amb_q[slot] = (uint32_t)qg; a.s.amb_row[slot] = (uint32_t)(row0 + c0 + j); }
            }
          }
        } else {
          if (qvalid) {
            float* dst = a.s.dense + (size_t)qg * a.s.ld + row0 + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (full_tile || c0 + j < ncols) dst[j] = sv[j];
          }
        }
      };

      // Warm start.  A list that begins at -inf lets the whole first tile through: ~3/4 of all the inserts a short
      // scan ever makes happen in its first tile and the MMA warp waits for the accumulator (C2: 16 k of 139 k cycles).
      // Before the first FULL tile of a part is filtered, its columns are cut into K groups and the smallest of the K
      // group maxima is taken: at least K of this tile's scores reach it, so everything BELOW it is outside the list's
      // top K whatever follows -- a valid first threshold for the price of one extra pass over the accumulator (no list,
      // no branches), and the rows it drops lie below the list's final last key like any other the list rejects.
      // (single-accumulator kernels with K = 8 / 16: in the others the extra live values spill the steady-state loop)
      if (mode == kModeTopk && !TWO && K <= 16 && cold && full_tile && warp_active && kHalfCols % K == 0 && kHalfCols / K <= 16) {
        const float v = warm_floor<K, TWO, kHalfCols>(acc0 + (uint32_t)cbeg, w0, w1);
        if (qvalid && v > -INFINITY && v < INFINITY) {
          // the next float below v (v is finite): rows EQUAL to v must still pass
          const uint32_t vb = __float_as_uint(v);
          const float below = v > 0.f ? __uint_as_float(vb - 1u) : (v < 0.f ? __uint_as_float(vb + 1u) : __uint_as_float(0x80000001u));
          tshare = fmaxf(tshare, below);
          if (tshare > thr) { thr = tshare; thr_raw = __fdiv_rd(thr, w0); }
        }
      }
      cold = false;
      // double-buffered TMEM reads: chunk i+1 is in flight while chunk i is processed
      uint32_t ra0[16], rb0[16], ra1[16], rb1[16];
      if (warp_active && nch > 0) load(cbeg, ra0, rb0);
      for (int i = 0; warp_active && i < nch; i += 2) {
        ptx::tmem_ld_wait();
        if (i + 1 < nch) load(cbeg + (i + 1) * 16, ra1, rb1);
        process(cbeg + i * 16, ra0, rb0);
        if (i + 1 < nch) {
          ptx::tmem_ld_wait();
          if (i + 2 < nch) load(cbeg + (i + 2) * 16, ra0, rb0);
          process(cbeg + (i + 1) * 16, ra1, rb1);
        }
      }
      // release this accumulator buffer to the (leader's) MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) ptx::mbar_arrive_remote(buf ? tempty1 : tempty0);
        else ptx::mbar_arrive(&tempty_bar[buf]);
      }
      ++it;
    })
    flush();
    if (dbg && warp == 2 && lane == 0) { a.dbg[blockIdx.x * 16 + 5] = w_tfull; a.dbg[blockIdx.x * 16 + 6] = t_fold; a.dbg[blockIdx.x * 16 + 7] = clock64() - t_begin; }
  }

#ifdef KEMR_DEBUG
  if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.dbg[blockIdx.x * 16 + 9] = (long long)t; }
#endif
#undef KEMR_FOR_TILES
  // no CTA of a pair may leave (or free TMEM) while its peer can still signal its barriers
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Encoding a tensor map is a driver call of ~1.5 us and a scan needs five; the map is a pure function of (address, rows,
// D, box), so the last few are kept per thread (a serving loop scans the same galleries from the same query buffer).
struct TmapCacheEntry { const void* base; int64_t rows; int D, box; int dev; CUtensorMap map; };
constexpr int kTmapCache = 16;

inline int make_tmap_2d_uncached(CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows);
inline int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows) {
  static thread_local TmapCacheEntry cache[kTmapCache];
  static thread_local int used = 0, next = 0;
  int dev = -1;
  cudaGetDevice(&dev);
  for (int i = 0; i < used; ++i) {
    const TmapCacheEntry& e = cache[i];
    if (e.base == base && e.rows == rows && e.D == D && e.box == box_rows && e.dev == dev) { *map = e.map; return 0; }
  }
  if (make_tmap_2d_uncached(map, base, rows, D, box_rows)) return 1;
  TmapCacheEntry& e = cache[next];
  e.base = base; e.rows = rows; e.D = D; e.box = box_rows; e.dev = dev; e.map = *map;
  next = (next + 1) % kTmapCache;
  if (used < kTmapCache) ++used;
  return 0;
}

inline int make_tmap_2d_uncached(CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { snprintf(g_mma_error, sizeof g_mma_error, "cuTensorMapEncodeTiled unavailable"); return 1; }
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const char* l2env = getenv("KEMR_TMAP_L2");                   // experiments: 0 none, 1 64 B, 2 128 B, 3 256 B (default)
  const CUtensorMapL2promotion promo = !l2env ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
      : (l2env[0] == '0' ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (l2env[0] == '1' ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
      : (l2env[0] == '2' ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B)));
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_mma_error, sizeof g_mma_error, "cuTensorMapEncodeTiled failed (%d) rows=%lld D=%d box=%d", (int)r,
             (long long)rows, D, box_rows);
    return 1;
  }
  return 0;
}

template <int K, int CL, bool TWO, int MODE>
inline int mma_launch_kpt(const CUtensorMap& mq, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& s0,
                          const CUtensorMap& s1, const MmaArgs& ma, const MmaPlan& pl, cudaStream_t st) {
  // the attribute sticks to the function on a device: set it once per (instantiation, device, size) of this thread
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_smem = 0;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && (dev != attr_dev || pl.smem + 1024 > attr_smem)) {
    e = cudaFuncSetAttribute(scan_mma_kernel<K, CL, TWO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem + 1024);
    if (e == cudaSuccess) { attr_dev = dev; attr_smem = pl.smem + 1024; }
  }
  if (e != cudaSuccess) { snprintf(g_mma_error, sizeof g_mma_error, "smem attribute: %s", cudaGetErrorString(e)); return 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl.ctas);
  cfg.blockDim = dim3(kMmaThreads);
  cfg.dynamicSmemBytes = pl.smem + 1024;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, scan_mma_kernel<K, CL, TWO, MODE>, mq, m0, m1, s0, s1, ma);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_mma_error, sizeof g_mma_error, "launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
template <int K, int MODE>
inline int mma_launch_km(const CUtensorMap& mq, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& s0,
                         const CUtensorMap& s1, const MmaArgs& ma, const MmaPlan& pl, cudaStream_t st) {
  if (pl.cl == 4) return pl.two ? mma_launch_kpt<K, 4, true, MODE>(mq, m0, m1, s0, s1, ma, pl, st) : mma_launch_kpt<K, 4, false, MODE>(mq, m0, m1, s0, s1, ma, pl, st);
  if (pl.cl == 2) return pl.two ? mma_launch_kpt<K, 2, true, MODE>(mq, m0, m1, s0, s1, ma, pl, st) : mma_launch_kpt<K, 2, false, MODE>(mq, m0, m1, s0, s1, ma, pl, st);
  return pl.two ? mma_launch_kpt<K, 1, true, MODE>(mq, m0, m1, s0, s1, ma, pl, st) : mma_launch_kpt<K, 1, false, MODE>(mq, m0, m1, s0, s1, ma, pl, st);
}
// the register lists exist in the top-k epilogue only: the count and dense kernels are instantiated once (K = 8)
inline int mma_launch_k(int K, const CUtensorMap& mq, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& s0,
                        const CUtensorMap& s1, const MmaArgs& ma, const MmaPlan& pl, cudaStream_t st) {
  if (ma.s.mode == kModeCount) return mma_launch_km<8, kModeCount>(mq, m0, m1, s0, s1, ma, pl, st);
  if (ma.s.mode == kModeDense) return mma_launch_km<8, kModeDense>(mq, m0, m1, s0, s1, ma, pl, st);
  switch (K) {
    case 8: return mma_launch_km<8, kModeTopk>(mq, m0, m1, s0, s1, ma, pl, st);
    case 16: return mma_launch_km<16, kModeTopk>(mq, m0, m1, s0, s1, ma, pl, st);
    default: return mma_launch_km<32, kModeTopk>(mq, m0, m1, s0, s1, ma, pl, st);
  }
}

// clusters of four CTAs (one 227 KB CTA per SM) the current device can hold at once
inline int mma_max_quads() {
  cudaFuncSetAttribute(scan_mma_kernel<8, 4, false, kModeTopk>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 64);
  cfg.blockDim = dim3(kMmaThreads);
  cfg.dynamicSmemBytes = kSmemBudget;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, scan_mma_kernel<8, 4, false, kModeTopk>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// the five tensor maps of one scan: queries, the galleries with the stage-sized box, the galleries with a 16-row box
// (partial tiles at the end of a unit's row range)
struct MmaMaps { CUtensorMap q, g0, g1, s0, s1; };

inline int mma_make_maps(const ScanArgs& s, const MmaPlan& pl, MmaMaps* m) {
  if (make_tmap_2d(&m->q, s.q, s.Q, s.D, pl.a_rows)) return 1;
  const int b_box = pl.cl == 4 ? 64 : (pl.pair ? 128 : pl.n_tile);     // gallery rows per TMA box
  if (make_tmap_2d(&m->g0, s.gal[0], s.M, s.D, b_box)) return 1;
  if (make_tmap_2d(&m->s0, s.gal[0], s.M, s.D, 16)) return 1;
  if (s.G > 1) {
    if (make_tmap_2d(&m->g1, s.gal[1], s.M, s.D, b_box)) return 1;
    if (make_tmap_2d(&m->s1, s.gal[1], s.M, s.D, 16)) return 1;
  } else { m->g1 = m->g0; m->s1 = m->s0; }
  return 0;
}

inline int mma_launch(const ScanArgs& s, const MmaPlan& pl, cudaStream_t st, const MmaMaps* cached = nullptr) {
  MmaMaps local;
  if (!cached) { if (mma_make_maps(s, pl, &local)) return 1; cached = &local; }
  const CUtensorMap &mq = cached->q, &m0 = cached->g0, &m1 = cached->g1, &ms0 = cached->s0, &ms1 = cached->s1;
  MmaArgs ma;
  ma.s = s;
  ma.n_tile = pl.n_tile; ma.n_qb = pl.n_qb; ma.n_t = pl.n_t; ma.stages = pl.stages; ma.kc = pl.kc;
  ma.merged = pl.merged; ma.ds = pl.ds; ma.kc_total = (pl.merged && !pl.ds) ? 2 * pl.kc : pl.kc;
  ma.a_rows = pl.a_rows; ma.parts = pl.parts; ma.q_pad = pl.q_pad; ma.q_blk = pl.q_blk; ma.gran = pl.gran; ma.vq = pl.vq;
  ma.split = make_split(s.M, pl.n_qb, pl.ctas / std::max(1, pl.cl), pl.gran);
  ma.dbg = nullptr;
  // Short lists (a thread sees few scores: survivor probability K/n per element stays high) append per lane without
  // warp votes; long scans keep the vote that skips chunks without survivors.  Measured: C1 233 -> 212 us, C2 101.5 -> 99 us.
  static const char* epi_env = getenv("KEMR_MMA_EPI");
  const long long units_run = pl.ctas / std::max(1, pl.cl);
  const double rtot = (double)pl.n_qb * (double)s.M;
  const double per_list = rtot / (double)std::max(1ll, units_run) / 2.0 / std::max(1, pl.vq);
  ma.epi_variant = epi_env ? atoi(epi_env) : (per_list < 8192.0 ? 2 : 0);
#ifdef KEMR_DEBUG
  static const bool debug = getenv("KEMR_MMA_DEBUG") != nullptr;
#else
  constexpr bool debug = false;
#endif
  if (debug) {
    static long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 1024 * 16 * sizeof(long long));
    cudaMemsetAsync(dbuf, 0, 1024 * 16 * sizeof(long long), st);
    ma.dbg = dbuf;
  }
  const int rc_launch = mma_launch_k(pl.K, mq, m0, m1, ms0, ms1, ma, pl, st);
  if (debug && rc_launch == 0) {
    static int printed = 0;
    cudaStreamSynchronize(st);
    if (printed++ < 4) {
      std::vector<long long> h((size_t)pl.ctas * 16);
      cudaMemcpy(h.data(), ma.dbg, h.size() * 8, cudaMemcpyDeviceToHost);
      double avg[16] = {0};
      const int stride = pl.pair ? 2 : 1;        // the MMA role runs in the leader CTA of every pair
      const int nu = pl.ctas / stride;
      for (int c = 0; c < pl.ctas; c += stride) for (int i = 0; i < 16; ++i) avg[i] += (double)h[(size_t)c * 16 + i] / nu;
      long long tmin = 1ll << 62, tmax = 0; int cmin = 0, cmax = 0;
      for (int c = 0; c < pl.ctas; c += stride) {
        const long long t = h[(size_t)c * 16 + 4];
        if (t < tmin) { tmin = t; cmin = c; }
        if (t > tmax) { tmax = t; cmax = c; }
      }
      fprintf(stderr, "[kemr mma dbg] cl=%d merged=%d vq=%d parts=%d n_tile=%d K=%d ctas=%d tile-equivalents/unit=%.2f stages=%d | producer: wait_empty=%.0f total=%.0f | mma: wait_full=%.0f wait_tempty=%.0f total=%.0f (min %lld @cta %d, max %lld @cta %d; max cta: wait_full=%lld wait_tempty=%lld) | epi(w2): wait_tfull=%.0f fold=%.0f total=%.0f  cycles\n",
              pl.cl, pl.merged, pl.vq, pl.parts, pl.n_tile, pl.K, pl.ctas, rtot / pl.n_tile / nu, pl.stages, avg[0], avg[1], avg[2], avg[3], avg[4],
              tmin, cmin, tmax, cmax, h[(size_t)cmax * 16 + 2], h[(size_t)cmax * 16 + 3], avg[5], avg[6], avg[7]);
      {
        long long s_min = 1ll << 62, s_max = 0, e_max = 0, e_min = 1ll << 62;
        for (int c = 0; c < pl.ctas; ++c) {
          const long long t0 = h[(size_t)c * 16 + 8], t1 = h[(size_t)c * 16 + 9];
          if (t0) { s_min = std::min(s_min, t0); s_max = std::max(s_max, t0); }
          if (t1) { e_max = std::max(e_max, t1); e_min = std::min(e_min, t1); }
        }
        fprintf(stderr, "[kemr mma dbg] globaltimer: first CTA start -> last CTA start %lld ns, -> first epilogue done %lld ns, -> last epilogue done %lld ns\n",
                s_max - s_min, e_min - s_min, e_max - s_min);
      }
      if (getenv("KEMR_MMA_DEBUG_TRACE")) {
        std::vector<long long> tr((size_t)8 * kTraceLen);
        cudaMemcpy(tr.data(), ma.dbg + kTraceBase, tr.size() * 8, cudaMemcpyDeviceToHost);
        const long long t0 = tr[0];
        fprintf(stderr, "[kemr mma trace] stage-iteration: cta0 producer saw-empty / issued | cta0 mma saw-full / committed | cta1 producer saw-empty / issued (cta1 on its own SM clock, relative to its first)\n");
        for (int i = 0; i < kTraceLen; ++i)
          fprintf(stderr, "[kemr mma trace] %3d: %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, tr[i] - t0, tr[kTraceLen + i] - t0,
                  tr[2 * kTraceLen + i] - t0, tr[3 * kTraceLen + i] - t0, tr[4 * kTraceLen + i] - tr[4 * kTraceLen], tr[5 * kTraceLen + i] - tr[4 * kTraceLen]);
      }
      if (getenv("KEMR_MMA_DEBUG_ALL")) {
        for (int c = 0; c < pl.ctas; c += stride)
          fprintf(stderr, "[kemr mma dbg] cta %d: mma total=%lld wait_full=%lld wait_tempty=%lld | epi total=%lld wait_tfull=%lld fold=%lld\n", c,
                  h[(size_t)c * 16 + 4], h[(size_t)c * 16 + 2], h[(size_t)c * 16 + 3], h[(size_t)c * 16 + 7], h[(size_t)c * 16 + 5], h[(size_t)c * 16 + 6]);
      }
    }
  }
  return rc_launch;
}

}  // namespace kemr
