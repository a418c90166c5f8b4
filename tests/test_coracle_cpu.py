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
