"""Property tests (hypothesis) of the mathematical claims the engine's design rests on, checked on the
CPU oracle: rank-by-counting == rank-by-sorting, the KG boost as a sparse side path, shard-and-merge == one scan,
row permutations permute indices."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O


def _scores(draw, q, m, ties):
    vals = draw(st.lists(st.integers(-40, 40) if ties else st.floats(-1, 1, allow_nan=False, width=32),
                         min_size=q * m, max_size=q * m))
    return (np.array(vals, dtype=np.float64) / (8.0 if ties else 1.0)).reshape(q, m)


@st.composite
def case(draw):
    q, m = draw(st.integers(1, 6)), draw(st.integers(1, 24))
    s = _scores(draw, q, m, draw(st.booleans()))
    tgt = np.array(draw(st.lists(st.integers(0, m - 1), min_size=q, max_size=q)), dtype=np.int64)
    return s, tgt


@settings(max_examples=200, deadline=None)
@given(case())
def test_rank_by_counting_equals_rank_by_stable_sort(c):
    s, tgt = c
    order = np.argsort(-s, axis=1, kind="stable")
    by_sort = np.array([int(np.where(order[i] == tgt[i])[0][0]) + 1 for i in range(len(tgt))])
    assert np.array_equal(O.canon_rank(s, tgt), by_sort)
    t = s[np.arange(len(tgt)), tgt][:, None]
    by_count = 1 + ((s > t) | ((s == t) & (np.arange(s.shape[1])[None, :] < tgt[:, None]))).sum(axis=1)
    assert np.array_equal(by_count, by_sort)


@settings(max_examples=200, deadline=None)
@given(case(), st.integers(1, 8), st.floats(0.05, 1.0), st.floats(0.0, 2.0), st.data())
def test_kg_boost_is_a_sparse_side_path(c, k, alpha, bonus, data):
    """final = alpha*clip + bonus*[hit], alpha > 0, bonus >= 0  =>  top-k(final) is inside top-k(clip) U hits."""
    s, _ = c
    q, m = s.shape
    k = min(k, m)
    hit = np.array(data.draw(st.lists(st.booleans(), min_size=q * m, max_size=q * m))).reshape(q, m)
    final = O.canon_fused64(s, None, 1.0, 0.0, alpha, np.where(hit, bonus, 0.0))
    top_final, _ = O.canon_topk(final, k)
    top_clip, _ = O.canon_topk(s, k)
    for i in range(q):
        allowed = set(top_clip[i].tolist()) | set(np.nonzero(hit[i])[0].tolist())
        assert set(top_final[i].tolist()) <= allowed


@settings(max_examples=150, deadline=None)
@given(case(), st.integers(1, 8), st.integers(1, 4))
def test_shard_and_merge_equals_one_scan(c, k, shards):
    s, _ = c
    q, m = s.shape
    k = min(k, m)
    want_i, want_s = O.canon_topk(s, k)
    bounds = [m * r // shards for r in range(shards + 1)]
    cand = [[] for _ in range(q)]
    for r in range(shards):
        lo, hi = bounds[r], bounds[r + 1]
        if hi == lo:
            continue
        idx, sc = O.canon_topk(s[:, lo:hi], min(k, hi - lo))
        for i in range(q):
            cand[i] += [(-sc[i, j], idx[i, j] + lo) for j in range(idx.shape[1])]
    for i in range(q):
        best = sorted(cand[i])[:k]                                  # (score desc, global index asc)
        assert [b[1] for b in best] == want_i[i].tolist() and [-b[0] for b in best] == want_s[i].tolist()


@settings(max_examples=150, deadline=None)
@given(case(), st.integers(1, 6), st.randoms(use_true_random=False))
def test_row_permutation_permutes_indices_when_scores_are_distinct(c, k, rnd):
    s, _ = c
    q, m = s.shape
    k = min(k, m)
    if any(len(set(row.tolist())) < m for row in s):
        return                                                      # ties are broken by index, which a permutation changes
    perm = list(range(m))
    rnd.shuffle(perm)
    perm = np.array(perm)
    a, _ = O.canon_topk(s, k)
    b, _ = O.canon_topk(s[:, perm], k)
    assert np.array_equal(perm[b], a)


def _rd_div32(a, b):
    """fp32 division rounded toward -inf (what __fdiv_rd does), from a binary64 quotient: the quotient of two 24-bit
    significands is either an fp32 value or at least ~2^-47 (relative) away from one, so the 2^-53 rounding of the
    binary64 division cannot move it across an fp32 value."""
    q = np.float64(a) / np.float64(b)
    f = np.float32(q)
    return np.nextafter(f, np.float32(-np.inf)) if np.float64(f) > q else f


@settings(max_examples=2000, deadline=None)
@given(st.floats(-4, 4, width=32, allow_nan=False), st.floats(2 ** -10, 4, width=32), st.floats(-8, 8, width=32, allow_nan=False),
       st.integers(-3, 3))
def test_raw_accumulator_compare_never_rejects_a_survivor(thr, w, raw, ulps):
    """Epilogue of the tcgen05 scan on short lists (csrc/scan_mma.cuh, append8_raw): the raw accumulator is compared
    with thr_raw = RD(thr / w) instead of weighting every score; `raw <= thr_raw  =>  fl(w * raw) <= thr` must hold
    for every fp32 thr, w > 0 and raw, so nothing that would pass the weighted test is ever skipped."""
    thr, w = np.float32(thr), np.float32(w)
    thr_raw = _rd_div32(thr, w)
    for r in (np.float32(raw), thr_raw):                             # a random value, and the neighbourhood of the cut
        for _ in range(abs(ulps)):
            r = np.nextafter(r, np.float32(np.inf if ulps > 0 else -np.inf))
        if r <= thr_raw:
            assert np.float32(w * r) <= thr


def _reglist_insert(sc, ix, s, r):
    """RegList<K>::insert of csrc/scan_mma.cuh, statement for statement (precondition: s > sc[K-1])."""
    K = len(sc)
    up_next = True
    for p in range(K - 1, 0, -1):
        up = s > sc[p - 1]
        nsc = sc[p - 1] if up else (s if up_next else sc[p])
        nix = ix[p - 1] if up else (r if up_next else ix[p])
        sc[p], ix[p] = nsc, nix
        up_next = up
    if up_next:
        sc[0], ix[0] = s, r


@settings(max_examples=500, deadline=None)
@given(st.sampled_from([1, 2, 8, 16]), st.lists(st.integers(-6, 6), min_size=0, max_size=60), st.integers(1, 16))
def test_register_list_keeps_the_exact_topk_with_lowest_row_first(K, values, trigger):
    """The epilogue's per-thread list: scores arrive in increasing row order, survivors (score > running threshold)
    are buffered and folded in batches (the threshold only tightens at a fold), ties keep the lower row ahead."""
    sc, ix = [-np.inf] * K, [0xffffffff] * K
    thr, buf = -np.inf, []
    for row, v in enumerate(values):
        s = v / 4.0
        if s > thr:
            buf.append((s, row))
        if len(buf) >= trigger or row == len(values) - 1:
            for bs, br in buf:                                       # fold: re-check against the tightened threshold
                if bs > thr:
                    _reglist_insert(sc, ix, bs, br)
                    thr = sc[K - 1]
            buf = []
    want = sorted(((-v / 4.0, row) for row, v in enumerate(values)))[:K]
    got = [(-s, r) for s, r in zip(sc, ix) if r != 0xffffffff]
    assert got == want
