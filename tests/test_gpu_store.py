"""The data path either side of the scan on the GPU: persisted gallery shards, device-built KG-hit CSR."""
import numpy as np
import pytest
import torch

from knowledge_enhanced_multimodal_retrieval_b200 import engine, fusion, store, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kg_set():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return synth.make_retrieval_set(Q=150, M=4000, D=256, seed=21, fused=True, lam=0.2, with_kg=True, diagonal=True)


def _csr_as_dicts(h: engine.KGHits):
    rp, col, bon = h.rowptr.cpu().numpy(), h.col.cpu().numpy(), h.bonus.cpu().numpy()
    return [list(zip(col[rp[i]:rp[i + 1]].tolist(), bon[rp[i]:rp[i + 1]].tolist())) for i in range(len(rp) - 1)]


@pytest.mark.parametrize("strategy,params", [("weighted", {"alpha": 0.8, "sparql_weight": 0.2}),
                                             ("additive", {"delta": 0.25}), ("adaptive", {"delta": 0.5})])
def test_device_csr_equals_host_builder(kg_set, strategy, params):
    s = kg_set
    # make the lists nastier: repeats, unknown uuids, URIs
    res = {k: list(v) for k, v in s.kg_results.items()}
    some = s.query_uuids[3]
    res[some] = res.get(some, []) + res.get(some, [])[:3] + ["http://x/not-there", f"http://y/z/{s.uuids[7]}", s.uuids[7]]
    lists = store.HitLists(store.IdMap(s.uuids), res, s.query_uuids)
    alpha_d, dev = lists.for_strategy(strategy, params)
    alpha_h, host = fusion.kg_hits_for_strategy(res, s.query_uuids, s.uuids, strategy, params)
    assert alpha_d == alpha_h and dev.max_per_query == host.max_per_query
    assert _csr_as_dicts(dev) == _csr_as_dicts(host)                    # same columns, same order, same binary64 bonus
    # a shard sees only its rows, re-based
    lo, hi = 1000, 2500
    _, sh = lists.for_strategy(strategy, params, lo, hi)
    assert _csr_as_dicts(sh) == _csr_as_dicts(host.shard(lo, hi))


def test_store_roundtrip_and_sharded_search(kg_set, tmp_path):
    s = kg_set
    st = store.EmbeddingStore.save(str(tmp_path / "gal"), s.image, s.target, s.uuids)
    assert (st.M, st.D, st.has_target) == (4000, 256, True)
    full = st.load()
    assert torch.equal(full.image, engine.quantize(s.image)) and torch.equal(full.target, engine.quantize(s.target))
    q = engine.quantize(s.query)
    lists = store.HitLists(st.idmap, s.kg_results, s.query_uuids)
    alpha, hits = lists.for_strategy("weighted", {"alpha": 0.8, "sparql_weight": 0.2})
    want_i, want_s = engine.scan_topk(q, full.image, full.target, 0.5, 0.5, alpha, hits, k=10)
    # three "ranks": each loads its byte range, builds its own CSR, emits global ids; merge == single scan
    parts_i, parts_s = [], []
    for r in range(3):
        shard = st.load(r, 3)
        lo = shard.idx_base
        _, h = lists.for_strategy("weighted", {"alpha": 0.8, "sparql_weight": 0.2}, lo, lo + shard.M)
        i, sc = shard.search(q, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=alpha, hits=h)
        parts_i.append(i); parts_s.append(sc)
    mi, ms = engine.merge_topk(torch.stack(parts_s), torch.stack(parts_i), 10)
    assert torch.equal(mi, want_i) and torch.equal(ms, want_s)
    # metrics through the store path == the mirror function on the raw arrays
    tidx = torch.arange(s.Q, device="cuda")
    ranks = engine.rank_targets(q, full.image, full.target, tidx, 0.5, 0.5, alpha, hits)
    got = fusion._metrics_from_ranks(ranks, [1, 5, 10, 20], True, True)
    want = fusion.evaluate_fused(s.query, s.target, s.image, s.kg_results, s.query_uuids, s.uuids, 0.5, 0.5, "weighted",
                                 {"alpha": 0.8, "sparql_weight": 0.2})
    assert {k: float(v) for k, v in got.items()} == {k: float(v) for k, v in want.items()}


def test_driver_sweep_equals_per_call_evaluations(kg_set):
    """evaluator.py:164-218 in one call == the same evaluations through the per-call mirror functions."""
    s = kg_set
    sub = slice(0, s.Q)
    img, tgt = s.image[: s.Q], s.target[: s.Q]                      # the driver's arrays are (N, D) with uuid_list for both
    uu = s.uuids[: s.Q]
    res = {uu[i]: list(s.kg_results[qu]) for i, qu in enumerate(s.query_uuids) if qu in s.kg_results}   # query uuid == artefact uuid
    got = fusion.evaluate_weight_and_alpha_sweep(s.query[sub], tgt, img, res, uu, weight_settings=((0.5, 0.5), (0.1, 0.9)),
                                                 alphas=(0.9, 0.5, 0.1))
    from knowledge_enhanced_multimodal_retrieval_b200 import metrics
    same = lambda a, b: {k: float(v) for k, v in a.items()} == {k: float(v) for k, v in b.items()}
    assert same(got["w0.5_0.5/T2I"], metrics.compute_retrieval_metrics(s.query[sub], img))
    assert same(got["w0.1_0.9/T2T"], metrics.compute_retrieval_metrics(s.query[sub], tgt))
    assert same(got["w0.1_0.9/Fused"], metrics.compute_retrieval_metrics_final(s.query[sub], tgt, img, t2i_weight=0.1, t2t_weight=0.9))
    for wi, wt in ((0.5, 0.5), (0.1, 0.9)):
        for a in (0.9, 0.5, 0.1):
            want = fusion.evaluate_fused(s.query[sub], tgt, img, res, uu, uu, wi, wt, "weighted",
                                         {"alpha": a, "sparql_weight": 1 - a})
            assert same(got[f"w{wi}_{wt}/alpha{a}"], want)
