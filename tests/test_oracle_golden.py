"""CPU: the numpy oracle reproduces every golden vector produced by the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import synth


def same(d_got, d_want):
    assert set(d_got) == set(d_want)
    for k in d_want:
        assert float(d_got[k]) == d_want[k], (k, float(d_got[k]), d_want[k])


def test_ref_metrics_small(golden, small_set):
    g = golden["metrics_small"]
    q, img, tgt, sim = (small_set[k] for k in ("query", "image", "target", "sim"))
    same(O.ref_retrieval_metrics(q, img, prefix="T2I"), g["retrieval_metrics_T2I"])
    same(O.ref_retrieval_metrics(q, tgt, k_values=[1, 3, 7]), g["retrieval_metrics_noprefix_k"])
    same(O.ref_retrieval_metrics_final(q, tgt, img), g["final_05_05"])
    same(O.ref_retrieval_metrics_final(q, tgt, img, prefix="F", t2i_weight=0.1, t2t_weight=0.9),
         g["final_01_09"])
    sq = [small_set[k] for k in ("sq_query", "sq_target", "sq_image")]
    same(O.ref_all_retrieval_metrics(*sq), g["all"])
    same(O.ref_all_retrieval_metrics(*sq, tasks=["T2I", "T2T"], compute_recall=False),
         g["all_T2I_T2T_mrr_only"])
    same(O.ref_all_retrieval_metrics(*sq, compute_recall=False), g["training"])
    # deprecated shims (metrics.py:285-352) = the dispatchers on (text, text, image) of the first variant
    sq_q, sq_t, sq_i = sq
    same(O.ref_all_retrieval_metrics(sq_t, sq_t, sq_i), g["shim_multi_mode"])
    same(O.ref_all_retrieval_metrics(sq_t, sq_t, sq_i, compute_recall=False), g["shim_single_4train"])
    same(O.ref_all_retrieval_metrics(sq_q, sq_q, sq_i, compute_recall=False), g["shim_multi_4train"])
    same(O.ref_recall_at_k(sim), g["recall_at_k_matrix"])
    same(O.ref_mrr_and_mean_rank(sim), g["mrr_matrix"])
    same(O.ref_metrics_from_matrix(sim, prefix="X"), g["metrics_fusion_matrix"])
    same(O.ref_metrics_from_matrix(sim), g["evaluate_retrieval"])


def test_ref_fusion_small_bit_exact(golden, small_set):
    sim = small_set["sim"]
    a = (sim, small_set["kg_results"], small_set["query_uuids"], small_set["uuids"])
    got = {
        "weighted_default": O.ref_weighted_fusion(*a),
        "weighted_09_01": O.ref_weighted_fusion(*a, alpha=0.9, sparql_weight=1 - 0.9),
        "weighted_renorm": O.ref_weighted_fusion(*a, alpha=0.6, sparql_weight=0.6),
        "additive_default": O.ref_additive_bonus_fusion(*a),
        "additive_013": O.ref_additive_bonus_fusion(*a, delta=0.13),
        "adaptive_default": O.ref_adaptive_additive_fusion(*a),
        "adaptive_custom": O.ref_adaptive_additive_fusion(*a, delta=0.3,
                                                          size_thresholds={2: 0.9, 10: 0.4, 25: 0.05}),
        "dispatch_weighted": O.ref_fuse_clip_and_text2sparql(*a, fusion_strategy="weighted",
                                                             fusion_params={"alpha": 0.4, "sparql_weight": 0.6}),
        "dispatch_additive": O.ref_fuse_clip_and_text2sparql(*a, fusion_strategy="additive"),
        "dispatch_adaptive": O.ref_fuse_clip_and_text2sparql(*a, fusion_strategy="adaptive",
                                                             fusion_params={"delta": 0.25}),
    }
    for name, mat in got.items():
        want = small_set["fusion_" + name]
        assert mat.dtype == np.float32 and mat.shape == want.shape
        assert np.array_equal(mat.view(np.uint32), want.view(np.uint32)), name
        same(O.ref_metrics_from_matrix(mat), golden["fusion_small_metrics"][name])
    with pytest.raises(ValueError):
        O.ref_fuse_clip_and_text2sparql(*a, fusion_strategy="nope")
    assert sim is a[0] and np.array_equal(sim, small_set["sim"])   # inputs never mutated


def test_ref_engine_list_fusion(golden):
    for c in golden["engine"]["fuse_cases"]:
        assert O.ref_fuse_clip_sparql_linear(c["clip"], c["sparql"], c["alpha"], c["beta"]) == c["out"]
    ci = golden["engine"]["call_inputs"]
    calls = golden["engine"]["calls"]
    fused = O.ref_fuse_clip_sparql_linear(ci["clip"], ci["sparql"], 0.8, 0.2)
    assert O.ref_threshold_filter(fused, 0) == calls["retrieve_text_default"]
    fused = O.ref_fuse_clip_sparql_linear(ci["clip"], ci["sparql"], 0.6, 0.4)
    assert O.ref_threshold_filter(fused, 0.3) == calls["retrieve_text_thr"]
    assert O.ref_threshold_filter(ci["clip"], 0) == calls["noknowledge_default"]
    assert O.ref_threshold_filter(ci["clip"], 0.25) == calls["noknowledge_thr"]


def test_ref_and_canonical_mid(golden):
    """Mid-size set regenerated from its seed: ref layer == golden; canonical layer == golden
    (no near-tie flips a rank for this seed)."""
    g = golden["metrics_mid"]
    s = synth.make_retrieval_set(Q=1500, M=1500, D=128, seed=23, fused=True, lam=0.4, with_kg=True)
    chk = [float(x.astype(np.float64).sum()) for x in (s.query, s.image, s.target)]
    assert chk == g["checksum"], "synthetic generator drifted from the committed fixtures"
    same(O.ref_all_retrieval_metrics(s.query, s.target, s.image), g["all"])
    si = O.canon_dot64(s.query, s.image)
    st = O.canon_dot64(s.query, s.target)
    r, c, _ = O.kg_hits_to_pairs(s.kg_results, s.query_uuids, s.uuids)
    for wi, wt in ((0.5, 0.5), (0.1, 0.9)):
        same(O.ref_retrieval_metrics_final(s.query, s.target, s.image, t2i_weight=wi, t2t_weight=wt),
             g[f"final_{wi}_{wt}"])
        clip = O.canon_fused64(si, st, wi, wt)
        same(O.metrics_from_ranks(O.canon_rank(clip, s.target_idx)), g[f"final_{wi}_{wt}"])
        for alpha in (0.9, 0.5, 0.1):
            bonus = O.canon_bonus_matrix(s.Q, s.M, r, c, 1 - alpha, dedupe=True)
            fused = O.canon_fused64(si, st, wi, wt, alpha=alpha, bonus=bonus)
            same(O.metrics_from_ranks(O.canon_rank(fused, s.target_idx)),
                 g[f"sweep_{wi}_{wt}_alpha{alpha}"])


def test_canonical_order_is_what_it_says():
    rng = np.random.default_rng(0)
    for D in (80, 264, 520):
        q = synth.round_to_bf16(rng.standard_normal((3, D), dtype=np.float32))
        g = synth.round_to_bf16(rng.standard_normal((5, D), dtype=np.float32))
        got = O.canon_dot64(q, g)
        for i in range(3):
            for j in range(5):
                run = [0.0] * 256                          # one running sum per position in a 256-block
                for d in range(D):
                    run[d % 256] += float(q[i, d]) * float(g[j, d])
                part = [((run[8 * l] + run[8 * l + 1]) + (run[8 * l + 2] + run[8 * l + 3])) +
                        ((run[8 * l + 4] + run[8 * l + 5]) + (run[8 * l + 6] + run[8 * l + 7])) for l in range(32)]
                n = 32
                while n > 1:
                    n //= 2
                    part = [part[l] + part[l + n] for l in range(n)]
                assert got[i, j] == part[0]
        assert np.allclose(got, q.astype(np.float64) @ g.astype(np.float64).T, rtol=0, atol=1e-12)


def test_canon_topk_and_rank_ties_nan():
    s = np.array([[0.5, 0.9, 0.9, np.nan, 0.1], [np.nan, 0.2, 0.2, 0.2, np.nan]])
    idx, val = O.canon_topk(s, 3)
    assert idx.tolist() == [[1, 2, 0], [1, 2, 3]]
    assert O.canon_rank(s, np.array([2, 4])).tolist() == [2, 5]
    assert O.canon_rank(s, np.array([3, 0])).tolist() == [5, 4]


def test_canonical_vs_reference_confined_to_near_ties(golden):
    """Where the fp32 reference and the binary64 contract disagree, it is only on rows the
    near-tie audit flags (the reference's own result there depends on its BLAS build)."""
    g = golden["metrics_mid_nearties"]["final_0.5_0.5"]
    s = synth.make_retrieval_set(Q=1500, M=1500, D=128, seed=21, fused=True, lam=0.25)
    clip = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), 0.5, 0.5)
    _, risky_rows = O.near_tie_audit(clip, s.target_idx, 20)
    got = O.metrics_from_ranks(O.canon_rank(clip, s.target_idx))
    assert 0 < risky_rows <= 32
    assert abs(float(got["Mean_Rank"]) - g["Mean_Rank"]) * s.Q <= risky_rows + 1e-6
    for k in (1, 5, 10, 20):
        assert abs(float(got[f"R@{k}"]) - g[f"R@{k}"]) * s.Q / 100.0 <= risky_rows + 1e-6


def test_grouped_and_gated_restatements_agree_with_the_canonical_forms():
    """The reference's grouped-ground-truth loops (baselines/evaluate_text_models.py:171-224) and gated-head formula
    (fusion_model.py:16-22) restated in the oracle, against the canonical (binary64, stable) forms on data without
    near ties: same ranks, hence same metrics."""
    from knowledge_enhanced_multimodal_retrieval_b200 import synth
    rng = np.random.default_rng(11)
    N, D = 60, 64
    base = synth.make_gallery(N, D, 5)
    cands = synth.round_to_bf16(synth.l2_normalize(np.repeat(base, 4, axis=0) * 0.7
                                                   + rng.normal(0, 1 / np.sqrt(D), (4 * N, D)).astype(np.float32)))
    query = synth.round_to_bf16(synth.l2_normalize(base * 0.6 + rng.normal(0, 1 / np.sqrt(D), (N, D)).astype(np.float32)))
    t2a = np.repeat(np.arange(N), 4)
    perm = rng.permutation(4 * N)
    cands, t2a = cands[perm], t2a[perm]
    ref = O.ref_grouped_metrics(query, cands, t2a)
    can = O.metrics_from_ranks(O.canon_grouped_rank(O.canon_dot64(query, cands), t2a), prefix="T2T")
    assert set(ref) == set(can)
    for k in ref:
        assert float(ref[k]) == pytest.approx(float(can[k]), abs=1e-9)
    gate = rng.uniform(0.1, 0.9, N).astype(np.float32)
    img, tgt = cands[:N], cands[N:2 * N]
    dense = O.ref_gated_scores(query, img, tgt, gate)
    canon = O.canon_fused64(O.canon_dot64(query, img), O.canon_dot64(query, tgt), gate.astype(np.float64),
                            (np.float32(1) - gate).astype(np.float64))
    assert np.abs(dense - canon).max() < 1e-5
    w = rng.normal(0, 0.3, D).astype(np.float32)
    g = O.ref_gate_linear(query, w, -0.5)
    assert g.dtype == np.float32 and np.allclose(g, 1 / (1 + np.exp(-(query.astype(np.float64) @ w.astype(np.float64) - 0.5))), atol=1e-6)


def test_fusion_heads_restatements_match_the_reference_modules(golden):
    """The oracle's restatements of the learned fusion heads against outputs of the UNMODIFIED torch modules of
    `src/clip/model/fusion_model.py` (tests/golden/make_golden.py, section "learned fusion heads")."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_set.npz"))
    from knowledge_enhanced_multimodal_retrieval_b200 import synth
    q, img, tgt = (synth.bf16_bits_to_f32(z[k]) for k in ("sq_query_bits", "sq_image_bits", "sq_target_bits"))
    H = golden["fusion_heads"]

    def check(name, scores, exact_rows=True):
        h = H[name]
        assert np.allclose(scores[0, :8], h["score_row0"], atol=2e-6), name
        assert abs(float(scores.astype(np.float64).sum()) - h["score_checksum"]) < 1e-2, name
        got = O.ref_metrics_from_matrix(scores)
        for k, v in h["metrics"].items():                        # at most one near-tie may flip between torch and numpy
            assert abs(float(got[k]) - v) <= 100.0 / len(scores) + 1e-9, (name, k, got[k], v)
        top5 = np.argsort(-scores, axis=1, kind="stable")[:, :5]
        assert (top5 == np.array(h["top5"])).mean() > 0.99, name

    p = H["simple_gated"]["params"]
    gate = O.ref_gate_linear(q, np.array(p["query_weight"], np.float32), p["bias"][0])
    assert np.abs(gate - np.array(H["simple_gated"]["gate"])).max() < 2e-6
    check("simple_gated", O.ref_gated_scores(q, img, tgt, gate))
    p = H["simple_gated_with_bias"]["params"]
    check("simple_gated_with_bias", O.ref_gated_scores(q, img, tgt, O.ref_gate_linear(q, np.array(p["query_weight"], np.float32), p["bias"][0])))
    p = H["gated_mlp"]["params"]
    D = q.shape[1]
    w1 = np.array(p["gate_net.0.weight"], np.float32).reshape(128, D)
    hid = np.maximum(q @ w1.T + np.array(p["gate_net.0.bias"], np.float32), 0)
    logit = hid @ np.array(p["gate_net.3.weight"], np.float32).reshape(-1) + np.float32(p["gate_net.3.bias"][0])
    check("gated_mlp", O.ref_gated_scores(q, img, tgt, (1 / (1 + np.exp(-logit))).astype(np.float32)))
    p = H["bilinear"]["params"]
    wi = np.array(p["W_image.weight"], np.float32).reshape(D, D)
    wt = np.array(p["W_target.weight"], np.float32).reshape(D, D)
    a = np.float32(1 / (1 + np.exp(-np.float32(p["alpha"][0]))))
    check("bilinear", a * (q @ (img @ wi.T).T) + (np.float32(1) - a) * (q @ (tgt @ wt.T).T))
    # the canonical (binary64, stable) form of the gated score gives the reference's metrics as well
    g64 = gate.astype(np.float64)
    can = O.canon_fused64(O.canon_dot64(q, img), O.canon_dot64(q, tgt), g64, (np.float32(1) - gate).astype(np.float64))
    want = H["simple_gated"]["metrics"]
    got = O.metrics_from_ranks(O.canon_rank(can, np.arange(len(q))))
    for k, v in want.items():
        assert abs(float(got[k]) - v) <= 100.0 / len(q) + 1e-9, (k, got[k], v)


def test_grouped_ground_truth_restatement_matches_the_reference_driver(golden):
    """`baselines/evaluate_text_models.py::evaluate_text_model` itself (unmodified, run with a fake encoder and a
    fake 5-variant dataset by make_golden.py), single and multi mode.  The driver re-normalises the embeddings in
    fp32 (:160-162) before scoring; with that step the oracle's restatement reproduces its numbers, and the
    canonical counting form on the bf16 inputs stays within one near-tie rank flip of them."""
    G = golden["grouped_text_models"]
    bf16 = [synth.bf16_bits_to_f32(np.array(v, dtype=np.uint16)) for v in G["variants_bits"]]
    renorm = [v / np.linalg.norm(v, axis=1, keepdims=True) for v in bf16]          # evaluate_text_models.py:161
    N = len(bf16[0])

    def pool(variants, exclude):
        cands = np.stack([variants[v][i] for i in range(N) for v in range(5) if v != exclude])
        return cands, np.repeat(np.arange(N), 4)

    def positions(query, cands, t2a):                                # the driver's loops (:237-279), fp32 + argsort
        sim = query @ cands.T
        return np.array([int(np.where(t2a[np.argsort(-sim[i])] == i)[0][0]) + 1 for i in range(N)])

    cands, t2a = pool(renorm, 0)
    ref = O.ref_grouped_metrics(renorm[0], cands, t2a)
    for k, v in G["single"].items():
        assert float(ref[k]) == pytest.approx(v, abs=1e-9), k
    multi_pos = np.concatenate([positions(renorm[qv], *pool(renorm, qv)) for qv in range(5)])
    multi = O.metrics_from_ranks(multi_pos, prefix="T2T")
    for k, v in G["multi"].items():
        assert float(multi[k]) == pytest.approx(v, abs=1e-9), k
    # canonical (binary64, stable) ranks on the bf16 inputs the engine sees
    can_pos = np.concatenate([O.canon_grouped_rank(O.canon_dot64(bf16[qv], pool(bf16, qv)[0]), pool(bf16, qv)[1])
                              for qv in range(5)])
    assert (can_pos != multi_pos).sum() <= 1                         # one fp32 near-tie at ranks 2/3 in this set
    can = O.metrics_from_ranks(can_pos, prefix="T2T")
    for k, v in G["multi"].items():
        assert abs(float(can[k]) - v) <= 100.0 / len(can_pos) + 1e-9, k
