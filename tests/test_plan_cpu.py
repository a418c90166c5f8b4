"""Property test of the tcgen05 scan's work plan (host logic, no GPU): the unit -> (query block, row range) cuts, the
tiles (full and partial) a unit walks and the part slots the epilogue derives from them must cover every row exactly
once and never collide.  The simulator below restates the index arithmetic of `scan_mma_kernel`
(csrc/scan_mma.cuh: make_split, unit_walk, unit_ordinal, KEMR_FOR_TILES, ord, slot)."""
import ctypes as C

import numpy as np
import pytest

from knowledge_enhanced_multimodal_retrieval_b200 import _lib

FIELDS = ("parts", "q_pad", "n_tile", "n_qb", "n_t", "ctas", "stages", "kc", "K", "cl", "gran", "vq", "all_slots",
          "two", "merged", "q_blk")


def plan(Q, M, D, G, k_sel, equal, sms=148, quads=33):
    out = (C.c_int64 * 16)()
    rc = _lib.load().kemr_debug_mma_plan(Q, M, D, G, k_sel, int(equal), sms, quads, out)
    return None if rc else dict(zip(FIELDS, list(out)))


def unit_begin(rtot, units, M, gran, u):
    x = rtot * u // units
    qb, r = divmod(x, M)
    return qb * M + r - r % gran


def unit_of(rtot, units, M, gran, pos):
    lo, hi = 0, units - 1
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if unit_begin(rtot, units, M, gran, mid) <= pos:
            lo = mid
        else:
            hi = mid - 1
    return lo


def make_split(M, n_qb, units, gran):
    return dict(M=M, n_qb=n_qb, units=units, gran=gran, s_full=units // n_qb)


def stripe_begin(w, s):
    if s >= w["s_full"] and w["units"] == w["s_full"] * w["n_qb"]:
        return w["M"]
    x = s * w["n_qb"] * w["M"] * 3 // (3 * w["units"] - 2 * (w["units"] - w["s_full"] * w["n_qb"]))
    return min(x - x % w["gran"], w["M"])


def unit_walk(w, u):
    """(base, mod, p_lo, p_hi, qb0) of unit u -- csrc/scan_mma.cuh::unit_walk."""
    full = w["s_full"] * w["n_qb"]
    if u < full:
        s = u // w["n_qb"]
        base = stripe_begin(w, s)
        ln = stripe_begin(w, s + 1) - base
        return base, max(ln, 1), 0, max(ln, 0), u % w["n_qb"]
    nr, j = w["units"] - full, u - full
    base = stripe_begin(w, w["s_full"])
    ln = w["M"] - base
    if ln <= 0:
        return base, 1, 0, 0, 0
    rtot = w["n_qb"] * ln
    return base, ln, unit_begin(rtot, nr, ln, w["gran"], j), unit_begin(rtot, nr, ln, w["gran"], j + 1), 0


def unit_ordinal(w, u, qb):
    full = w["s_full"] * w["n_qb"]
    if u < full:
        return u // w["n_qb"]
    nr = w["units"] - full
    base = stripe_begin(w, w["s_full"])
    ln = w["M"] - base
    first = unit_of(w["n_qb"] * ln, nr, ln, w["gran"], qb * ln)
    return w["s_full"] + (u - full - first)


def seg_tile_rows(rows, n_tile, gran):
    if gran >= n_tile or rows <= 0:
        return n_tile
    nt = -(-rows // n_tile)
    return min(n_tile, -(-(-(-rows // nt)) // gran) * gran)


def simulate(p, M):
    """(covered[qb, row] count, owner[(qb, ord)] -> unit, MMA rows per unit) exactly as the kernel's roles walk a unit's
    share of the work (unit_walk, KEMR_FOR_TILES) and as its epilogue derives the part slots."""
    units = p["ctas"] // p["cl"]
    n_tile, n_qb, gran, vq = p["n_tile"], p["n_qb"], p["gran"], p["vq"]
    w = make_split(M, n_qb, units, gran)
    covered = np.zeros((n_qb, M), dtype=np.int32)
    owner = {}
    work = []
    partial_ok = (not p["two"]) and p["cl"] != 4 and gran < n_tile
    for unit in range(units):
        base, mod, p_lo, p_hi, qb0 = unit_walk(w, unit)
        assert p_lo <= p_hi
        mma_rows = 0
        pos = p_lo
        while pos < p_hi:
            b = pos // mod
            qb = qb0 + b
            assert qb < n_qb, "unit reaches past the last query block"
            blk0 = b * mod
            rend = min(p_hi, blk0 + mod) - blk0
            ordinal = unit_ordinal(w, unit, qb)
            r = pos - blk0
            tsz = seg_tile_rows(rend - r, n_tile, gran if partial_ok else n_tile)
            while r < rend:
                row0 = base + r
                ncols = min(tsz, rend - r)
                assert row0 + ncols <= M
                g_ = 32 if p["cl"] >= 2 else 16
                nmma = min(256, (ncols + g_ - 1) // g_ * g_) if partial_ok else 256
                assert nmma % g_ == 0 and g_ <= nmma <= 256 and (p["two"] or nmma >= ncols)
                mma_rows += nmma if not p["two"] else n_tile
                covered[qb, row0:row0 + ncols] += 1
                o = (row0 * vq // M if vq > 1 else 0) + ordinal
                assert 0 <= 2 * o + 1 < p["parts"], f"slot {2 * o + 1} outside the {p['parts']} parts"
                assert owner.setdefault((qb, o), unit) == unit, "two units write the same part slot"
                r += tsz
            pos = blk0 + rend
        work.append(mma_rows)
    # the units of one full stripe scan the same rows (that is what keeps the gallery reads of the blocks together)
    for s_ in range(w["s_full"]):
        walks = {unit_walk(w, s_ * n_qb + qb)[:4] for qb in range(n_qb)}
        assert len(walks) == 1
    return covered, owner, work


SHAPES = [
    # Q, M, D, G, k_sel, equal
    (1000, 43000, 768, 2, 16, True), (4300, 43000, 512, 1, 16, False), (4096, 1_250_000, 768, 1, 112, False),
    (8192, 12_500_000, 768, 1, 16, False), (64, 2_000_000, 768, 1, 16, False), (3300, 3400, 64, 1, 16, False),
    (2700, 3000, 128, 2, 16, False), (5, 300, 8, 1, 16, False), (129, 1, 64, 1, 16, False), (257, 300, 128, 2, 16, True),
    (700, 9000, 256, 1, 112, False), (33000, 5000, 64, 1, 16, False), (520, 7000, 256, 1, 112, False),
]


@pytest.mark.parametrize("sms,quads", [(148, 33), (148, 0), (132, 30), (2, 0)])
@pytest.mark.parametrize("shape", SHAPES)
def test_plan_tiles_the_work_without_slot_collisions(shape, sms, quads):
    p = plan(*shape, sms=sms, quads=quads)
    if p is None:
        pytest.skip("shape not plannable on this device size (the API falls back to the warp-dot kernel)")
    Q = shape[0]
    assert p["q_pad"] % p["q_blk"] == 0 and p["q_pad"] >= Q and p["n_qb"] * p["q_blk"] == p["q_pad"]
    assert p["ctas"] % p["cl"] == 0 and 1 <= p["ctas"] <= sms and (p["cl"] != 4 or p["ctas"] // 4 <= quads)
    assert p["stages"] >= 2 and p["K"] in (8, 16, 32) and 2 <= p["parts"] <= 304 and p["parts"] % 2 == 0
    assert p["n_t"] == -(-shape[1] // p["n_tile"]) and p["kc"] == -(-shape[2] // 64)
    covered, owner, work = simulate(p, shape[1])
    assert (covered == 1).all(), "every (query block, gallery row) must be scanned exactly once"
    if p["all_slots"]:
        assert len(owner) == p["n_qb"] * (p["parts"] // 2), "all_slots promises that no part slot stays unwritten"
    if p["gran"] < p["n_tile"] and shape[1] >= 4 * p["n_tile"]:
        # partial tiles level the tensor work: no unit does more than the mean plus the rounding of its segments (a
        # remainder unit may walk every query block: one 32-row round-up and one snapped boundary per block)
        # (remainder units count as half a unit, so the full units carry up to units / (units - r / 2) of the mean)
        assert max(work) <= sum(work) / len(work) * 1.12 + 32 * (2 * p["n_qb"] + 4), (max(work), sum(work) / len(work))


def test_plan_random_shapes():
    rng = np.random.default_rng(0)
    n = 0
    for _ in range(400):
        Q = int(rng.choice([rng.integers(5, 300), rng.integers(300, 9000)]))
        M = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 3_000_000)]))
        D = int(rng.choice([8, 64, 128, 512, 768, 1024]))
        G = int(rng.integers(1, 3))
        k_sel = int(rng.choice([16, 24, 56, 112, 128]))
        equal = bool(rng.integers(0, 2)) and G == 2
        sms, quads = [(148, 33), (148, 0), (132, 30), (20, 4)][int(rng.integers(0, 4))]
        p = plan(Q, M, D, G, k_sel, equal, sms, quads)
        if p is None:
            continue
        n += 1
        if p["n_qb"] * M > 40_000_000:
            continue                                   # the row-level simulation of huge shapes is covered by SHAPES
        covered, owner, _ = simulate(p, M)
        assert (covered == 1).all(), (Q, M, D, G, k_sel, equal, sms, quads, p)
        if p["all_slots"]:
            assert len(owner) == p["n_qb"] * (p["parts"] // 2), (Q, M, D, G, k_sel, p)
    assert n > 200


def test_short_lists_with_virtual_parts_are_preferred_when_few_units_share_a_block():
    """With the per-query shared threshold a restarted list starts warm, so the plan's cost model takes lists of 8 cut
    into virtual parts over lists of 16 (C1: 4300 queries x 43 000 x 512-d, 17 query blocks on 74 pairs -- measured
    181 us against 199 us); where many units already share a block (C2) nothing changes."""
    c1 = plan(4300, 43000, 512, 1, 16, False)
    assert c1["K"] == 8 and c1["vq"] > 1 and c1["parts"] <= 304 and c1["all_slots"] == 0
    c2 = plan(1000, 43000, 768, 2, 16, True)
    assert (c2["K"], c2["vq"], c2["parts"], c2["merged"]) == (8, 1, 38, 1)
    top100 = plan(4096, 1250000, 768, 1, 112, False)
    assert top100["vq"] > 1 and top100["cl"] == 4
