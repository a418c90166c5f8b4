"""Property test of the tcgen05 scan's work plan (host logic, no GPU): the unit -> (query block, gallery tile) ranges
and the part slots the epilogue derives from them must tile the work exactly and never collide.  The simulator below
restates the index arithmetic of `scan_mma_kernel` (csrc/scan_mma.cuh: w_lo/w_hi, c_first, ord, slot)."""
import ctypes as C

import numpy as np
import pytest

from knowledge_enhanced_multimodal_retrieval_b200 import _lib

FIELDS = ("parts", "q_pad", "n_tile", "n_qb", "n_t", "ctas", "stages", "kc", "K", "cl", "upq", "vq", "all_slots",
          "two", "merged", "q_blk")


def plan(Q, M, D, G, k_sel, equal, sms=148, quads=33):
    out = (C.c_int64 * 16)()
    rc = _lib.load().kemr_debug_mma_plan(Q, M, D, G, k_sel, int(equal), sms, quads, out)
    return None if rc else dict(zip(FIELDS, list(out)))


def simulate(p):
    """(covered[(qb, t)] count, owner[(qb, ord)] -> unit) exactly as the kernel's producer / epilogue index them."""
    units = p["ctas"] // p["cl"]
    n_t, n_qb, upq, vq = p["n_t"], p["n_qb"], p["upq"], p["vq"]
    W = n_qb * n_t
    covered = np.zeros((n_qb, n_t), dtype=np.int32)
    owner = {}
    for unit in range(units):
        if upq > 0:
            j, base = unit % upq, (unit // upq) * n_t
            w_lo, w_hi = base + n_t * j // upq, base + n_t * (j + 1) // upq
        else:
            w_lo, w_hi = W * unit // units, W * (unit + 1) // units
        w = w_lo
        while w < w_hi:
            qb = w // n_t
            assert qb < n_qb, "unit reaches past the last query block"
            c_first = qb * upq if upq > 0 else ((qb * n_t + 1) * units - 1) // W
            t_end = min(w_hi, (qb + 1) * n_t)
            ts = np.arange(w - qb * n_t, t_end - qb * n_t)
            covered[qb, ts] += 1
            ords = (ts * vq // n_t if vq > 1 else np.zeros_like(ts)) + (unit - c_first)
            for o in np.unique(ords):
                assert 0 <= 2 * int(o) + 1 < p["parts"], f"slot {2 * int(o) + 1} outside the {p['parts']} parts"
                assert owner.setdefault((qb, int(o)), unit) == unit, "two units write the same part slot"
            w = t_end
    return covered, owner


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
    covered, owner = simulate(p)
    assert (covered == 1).all(), "every (query block, gallery tile) must be scanned exactly once"
    if p["all_slots"]:
        assert len(owner) == p["n_qb"] * (p["parts"] // 2), "all_slots promises that no part slot stays unwritten"


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
        covered, owner = simulate(p)
        assert (covered == 1).all(), (Q, M, D, G, k_sel, equal, sms, quads, p)
        if p["all_slots"]:
            assert len(owner) == p["n_qb"] * (p["parts"] // 2), (Q, M, D, G, k_sel, p)
    assert n > 200
