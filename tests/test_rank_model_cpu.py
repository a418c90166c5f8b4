"""Soundness of the rank path (kemr_rank_count: csrc/select.cuh rank_band / rank_amb / rank_hits kernels around the
counting epilogue of the scan), checked on a CPU restatement with hypothesis: for ANY fp32 scan scores within eps of
the canonical ones and any KG hits, 1 + count is the exact canonical rank of the target (score descending, ties to
the lowest index) -- the quantity the reference gets from two full-row argsorts (metrics.py:34,62,68)."""
import numpy as np
from hypothesis import given, settings, strategies as st


def ahead(sa, ia, sb, ib):
    return sa > sb or (sa == sb and ia < ib)


def rank_model(score32, canon_clip, alpha, hits, target, eps):
    """score32: fp32 clip-level scan scores [M]; canon_clip: canonical clip-level scores; hits: row -> bonus."""
    t = alpha * canon_clip[target] + hits.get(target, 0.0)          # canonical final score of the target (score_pairs)
    c = t / alpha
    e = eps * (1.0 + 1.0 / 64.0) + abs(c) * 1e-12
    lo = np.nextafter(np.float32(c - e), np.float32(-np.inf)) if np.float32(c - e) > c - e else np.float32(c - e)
    hi = np.nextafter(np.float32(c + e), np.float32(np.inf)) if np.float32(c + e) < c + e else np.float32(c + e)
    count, amb = 0, []
    for j, s in enumerate(score32):                                  # the scan's counting epilogue
        if s > hi:
            count += 1
        elif s >= lo:
            amb.append(j)
    for j in amb:                                                    # rank_amb_kernel: exact, bonus ignored
        if ahead(alpha * canon_clip[j], j, t, target):
            count += 1
    for j, b in hits.items():                                        # rank_hits_kernel: replace the unboosted verdict
        f0 = alpha * canon_clip[j]
        count += int(ahead(f0 + b, j, t, target)) - int(ahead(f0, j, t, target))
    return 1 + count


@st.composite
def scenario(draw):
    M = draw(st.integers(1, 80))
    eps = 2e-5
    grid = draw(st.sampled_from([1e-5, 3e-5, 1e-3, 0.05]))
    base = np.array(draw(st.lists(st.integers(-50, 50), min_size=M, max_size=M)), dtype=np.float64) * grid
    noise = np.array(draw(st.lists(st.floats(-1, 1), min_size=M, max_size=M)))
    score32 = base.astype(np.float32)
    canon = score32.astype(np.float64) + noise * eps * 0.999
    if draw(st.booleans()):
        canon = score32.astype(np.float64)                           # exact ties between fp32 and canonical
    alpha = draw(st.sampled_from([1.0, 0.8, 0.3]))
    nh = draw(st.integers(0, min(M, 6)))
    rows = draw(st.lists(st.integers(0, M - 1), min_size=nh, max_size=nh, unique=True))
    bonus = draw(st.sampled_from([0.0, 1e-5, 0.2, 0.7]))
    target = draw(st.integers(0, M - 1))
    return score32, canon, alpha, {r: bonus for r in rows}, target, eps


@settings(max_examples=800, deadline=None)
@given(scenario())
def test_counting_rank_is_the_exact_canonical_rank(sc):
    score32, canon, alpha, hits, target, eps = sc
    got = rank_model(score32, canon, alpha, hits, target, eps)
    final = [alpha * canon[j] + hits.get(j, 0.0) for j in range(len(canon))]
    want = 1 + sum(ahead(final[j], j, final[target], target) for j in range(len(canon)))
    assert got == want
