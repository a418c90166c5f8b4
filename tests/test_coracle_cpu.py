"""CPU: the C restatement (oracle/kemr_oracle.c) agrees bit-for-bit with the numpy oracle."""
import numpy as np

from oracle import coracle as CO
from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import engine, synth


def test_c_oracle_equals_numpy_oracle():
    for D in (8, 72, 200, 768):
        s = synth.make_retrieval_set(Q=23, M=311, D=D, seed=40 + D, fused=True, lam=0.3, with_kg=True, diagonal=False)
        si, st = O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target)
        assert np.array_equal(CO.scores(s.query, s.image), O.canon_fused64(si, None))
        want = O.canon_fused64(si, st, 0.1, 0.9, 0.7)
        assert np.array_equal(CO.scores(s.query, s.image, s.target, 0.1, 0.9, 0.7), want)
        idx, sc, rk = CO.topk_rank(s.query, s.image, s.target, 0.1, 0.9, 0.7, k=20, target=s.target_idx)
        widx, wsc = O.canon_topk(want, 20)
        assert np.array_equal(idx, widx) and np.array_equal(sc, wsc)
        assert np.array_equal(rk, O.canon_rank(want, s.target_idx))
        # with a KG-hit CSR (unique columns, aggregated bonus)
        cols, _ = engine.kg_pairs(s.kg_results, s.query_uuids, s.uuids)
        cols = [list(dict.fromkeys(c)) for c in cols]
        rp = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int64)
        col = np.array([j for c in cols for j in c], np.int32)
        bon = np.full(len(col), 0.3)
        r, c, _ = O.kg_hits_to_pairs(s.kg_results, s.query_uuids, s.uuids)
        want = O.canon_fused64(si, st, 0.5, 0.5, 0.7, O.canon_bonus_matrix(s.Q, s.M, r, c, 0.3, dedupe=True))
        assert np.array_equal(CO.scores(s.query, s.image, s.target, 0.5, 0.5, 0.7, (rp, col, bon)), want)


def test_c_oracle_short_gallery_padding():
    s = synth.make_retrieval_set(Q=3, M=4, D=16, seed=1, fused=False, diagonal=False)
    idx, sc, _ = CO.topk_rank(s.query, s.image, k=7)
    assert (idx[:, 4:] == -1).all() and np.isneginf(sc[:, 4:]).all()


def test_c_oracle_per_query_weights_equal_numpy_oracle():
    """Gated fusion heads: per-query (w_a, w_b) arrays through the C oracle == oracle.canon_fused64 with arrays."""
    import numpy as np
    from oracle import oracle as O
    from knowledge_enhanced_multimodal_retrieval_b200 import synth
    s = synth.make_retrieval_set(Q=40, M=300, D=128, seed=13, fused=True, lam=0.2, diagonal=True)
    rng = np.random.default_rng(1)
    gate = rng.uniform(0.05, 0.95, 40).astype(np.float32)
    wa, wb = gate.astype(np.float64), (np.float32(1) - gate).astype(np.float64)
    want = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), wa, wb)
    got = CO.scores(s.query, s.image, s.target, wa, wb)
    assert np.array_equal(got, want)
    idx, sc, rank = CO.topk_rank(s.query, s.image, s.target, wa, wb, k=7, target=s.target_idx)
    widx, wsc = O.canon_topk(want, 7)
    assert np.array_equal(idx, widx) and np.array_equal(sc, wsc) and np.array_equal(rank, O.canon_rank(want, s.target_idx))
    again = CO.scores(s.query, s.image, s.target, 0.5, 0.5)            # the scalars are back after the call
    assert np.array_equal(again, O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), 0.5, 0.5))
