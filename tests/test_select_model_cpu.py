"""Soundness of the select stage's pruning and certificate (csrc/select.cuh, stages A1-E), checked on a CPU restatement
of its logic with hypothesis: whenever the certificate bit stays clear, the returned top-k is the exact top-k of the
whole gallery under the canonical scores -- for ANY fp32 scan scores within eps of the canonical ones, any split of
the rows into part lists, any KG hits.  (The CUDA kernel itself is checked against the oracle by the GPU tests; this
pins the reasoning the kernel implements, including the rare branches: fewer lists than K, short lists, the fp32
pre-screen past the k-th candidate.)"""
import numpy as np
from hypothesis import given, settings, strategies as st


def key(score32, row):
    return (float(score32), -int(row))            # larger tuple = higher fp32 score, then lower row; None = empty slot


def select_model(lists, Kp, K, k, eps, alpha, canon, hits):
    """lists: P part lists, each a descending list of up to Kp keys (exact top-Kp of its rows by fp32 key).
    Returns (rows of the top-k, flag)."""
    bound = None

    def reject(x):
        nonlocal bound
        if x is not None and (bound is None or x > bound):
            bound = x

    # A1: the (at most K) lists with the largest heads
    heads = sorted(((l[0], i) for i, l in enumerate(lists) if l), reverse=True)
    sel = [i for _, i in heads[:K]]
    for h, _ in heads[K:]:
        reject(h)
    nlist = len(sel)
    # lower bound of the K-th best key
    cut = None
    if nlist >= K:
        cut = lists[sel[K - 1]][0]
    elif nlist > 0:
        jc = min(Kp, -(-K // nlist))
        sub = sorted((x for i in sel for x in lists[i][:jc]), reverse=True)
        if len(sub) >= K:
            cut = sub[K - 1]
    # A2: survivors, ranked
    surv = []
    for i in sel:
        for j, x in enumerate(lists[i]):
            if cut is None or x >= cut:
                surv.append(x)
            else:
                reject(x)
            if j == Kp - 1:
                reject(x)                         # a full part list rejected rows below its last key
    surv.sort(reverse=True)
    selk = surv[:K]
    for x in surv[K:]:
        reject(x)
    # A3: fp32 pre-screen past the k-th candidate
    nsel = len(selk)
    if nsel > k:
        kth32 = selk[k - 1][0] - 2.0 * eps * (1.0 + 1.0 / 64.0)
        for i in range(k, nsel):
            if selk[i][0] < kth32:
                reject(selk[i])
                nsel = i
                break
    # B-D
    cand = {-x[1]: 0.0 for x in selk[:nsel]}
    for row, bonus in hits.items():
        cand[row] = bonus
    final = sorted(((-(alpha * canon[r] + b), r) for r, b in cand.items()))
    out = [r for _, r in final[:k]]
    # E
    flag = 0
    if bound is not None:
        reach = alpha * (bound[0] + eps * (1.0 + 1.0 / 64.0)) + 1e-300
        if not (len(final) >= k and -final[k - 1][0] > reach):
            flag = 1
    return out, flag


@st.composite
def scenario(draw):
    M = draw(st.integers(1, 120))
    P = draw(st.integers(1, 10))
    Kp = draw(st.sampled_from([1, 2, 4, 8]))
    K = draw(st.integers(1, 16))
    k = draw(st.integers(1, K))
    eps = 2e-5
    grid = draw(st.sampled_from([1e-5, 4e-5, 1e-3, 0.05]))          # coarse grids make fp32 near-ties and exact ties
    base = np.array(draw(st.lists(st.integers(-60, 60), min_size=M, max_size=M)), dtype=np.float64) * grid
    noise = np.array(draw(st.lists(st.floats(-1, 1), min_size=M, max_size=M)))
    canon = base + noise * eps * 0.999                               # |fp32 - canonical| <= eps, the kernel's assumption
    score32 = base.astype(np.float32)
    part = np.array(draw(st.lists(st.integers(0, P - 1), min_size=M, max_size=M)))
    alpha = draw(st.sampled_from([1.0, 0.8, 0.3]))
    nh = draw(st.integers(0, min(M, 6)))
    hit_rows = draw(st.lists(st.integers(0, M - 1), min_size=nh, max_size=nh, unique=True))
    bonus = draw(st.sampled_from([0.0, 0.2, 0.7]))
    return M, P, Kp, K, k, eps, canon, score32, part, alpha, {r: bonus for r in hit_rows}


@settings(max_examples=600, deadline=None)
@given(scenario())
def test_clear_certificate_implies_exact_topk(sc):
    M, P, Kp, K, k, eps, canon, score32, part, alpha, hits = sc
    lists = []
    for p in range(P):
        rows = np.nonzero(part == p)[0]
        keys = sorted((key(score32[r], r) for r in rows), reverse=True)[:Kp]
        lists.append(keys)
    out, flag = select_model(lists, Kp, K, k, eps, alpha, canon, hits)
    truth = sorted((-(alpha * canon[r] + hits.get(r, 0.0)), r) for r in range(M))[:k]
    if flag == 0:
        assert out == [r for _, r in truth][:len(out)], (out, truth)
        if M >= k:
            assert len(out) == k


def test_prescreen_margin_keeps_near_ties_of_the_kth_candidate():
    """The fp32 pre-screen (A3) may only drop candidates more than 2 eps below the k-th fp32 score: here the fp32
    runner-up (1e-5 below) is the canonical winner, nothing else is rejected, and the certificate stays clear."""
    eps = 2e-5
    score32 = np.array([0.50000, 0.49999, 0.1], dtype=np.float32)
    canon = np.array([0.5 - 1.9e-5, 0.49999 + 1.9e-5, 0.1])
    lists = [sorted((key(score32[r], r) for r in range(3)), reverse=True)]
    out, flag = select_model(lists, Kp=8, K=3, k=1, eps=eps, alpha=1.0, canon=canon, hits={})
    assert (out, flag) == ([1], 0)
    # a candidate well below the margin is dropped, counted as rejected, and the result is still certified
    score32 = np.array([0.5, 0.3, 0.1], dtype=np.float32)
    canon = score32.astype(np.float64)
    lists = [sorted((key(score32[r], r) for r in range(3)), reverse=True)]
    out, flag = select_model(lists, Kp=8, K=3, k=1, eps=eps, alpha=1.0, canon=canon, hits={})
    assert (out, flag) == ([0], 0)
    # the dropped candidate comes back through the hit list when the knowledge graph boosts it past the leader
    out, flag = select_model(lists, Kp=8, K=3, k=1, eps=eps, alpha=1.0, canon=canon, hits={2: 0.7})
    assert (out, flag) == ([2], 0)


@st.composite
def shared_threshold_scenario(draw):
    sc = draw(scenario())
    M = sc[0]
    order = draw(st.permutations(list(range(M))))
    stale = draw(st.lists(st.integers(0, 3), min_size=M, max_size=M))      # how many publications a reader lags behind
    return sc, order, stale


@settings(max_examples=600, deadline=None)
@given(shared_threshold_scenario())
def test_shared_threshold_between_lists_keeps_the_certificate_sound(arg):
    """The tcgen05 epilogue prunes every list of a query with the largest K-th score any FULL list has published so
    far (scan_mma.cuh: thr_pub), read with arbitrary staleness, rows of different lists arriving in any interleaving.
    A row dropped that way has Kp better rows in one list whose final last key is at least the threshold used, and the
    select stage puts every full list's last key into its rejection bound -- so a clear certificate must still imply
    the exact top-k, although the lists are no longer the exact top-Kp of their rows."""
    (M, P, Kp, K, k, eps, canon, score32, part, alpha, hits), order, stale = arg
    lists = [[] for _ in range(P)]
    published = [float("-inf")]                                   # history of the shared threshold (monotone)
    for n, r in enumerate(order):
        seen = published[max(0, len(published) - 1 - stale[n])]   # a possibly stale view
        p = int(part[r])
        own = lists[p][-1][0] if len(lists[p]) == Kp else float("-inf")
        if not float(score32[r]) > max(seen, own):                # strict compare on the fp32 score, as the kernel does
            continue
        lists[p] = sorted(lists[p] + [key(score32[r], r)], reverse=True)[:Kp]
        if len(lists[p]) == Kp:
            published.append(max(published[-1], lists[p][-1][0]))
    out, flag = select_model(lists, Kp, K, k, eps, alpha, canon, hits)
    truth = sorted((-(alpha * canon[r] + hits.get(r, 0.0)), r) for r in range(M))[:k]
    if flag == 0:
        assert out == [r for _, r in truth][:len(out)], (out, truth)
        if M >= k:
            assert len(out) == k


@settings(max_examples=400, deadline=None)
@given(st.sampled_from([(8, 128), (16, 128), (8, 64), (16, 64)]),
       st.lists(st.integers(-2000, 2000), min_size=128, max_size=128), st.sampled_from([1e-4, 1e-3, 0.25]))
def test_warm_start_floor_never_drops_a_row_of_the_lists_top_k(shape, ints, grid):
    """scan_mma.cuh warm start: the first full tile's half of HC columns is cut into K groups and the smallest of the K
    group maxima, v, becomes the list's first threshold (everything <= the float just below v is dropped).  At least K
    of the tile's scores are >= v, so no dropped score can belong to the top K of the rows the list goes on to see."""
    K, HC = shape
    scores = (np.array(ints[:HC], dtype=np.float64) * grid).astype(np.float32)
    g = HC // K
    v = min(scores[j * g:(j + 1) * g].max() for j in range(K))
    floor = np.nextafter(np.float32(v), np.float32(-np.inf))
    kept = scores[scores > floor]                                  # the kernel's strict compare against the floor
    assert kept.size >= K and (scores >= v).sum() >= K
    kth = np.sort(scores)[::-1][K - 1]                             # K-th largest of the tile (with multiplicity)
    assert floor < kth                                             # so every top-K score of the tile passes ...
    assert np.all(scores[scores <= floor] < kth)                   # ... and every dropped one is strictly below all of them
