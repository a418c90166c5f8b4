"""GPU parity: every CUDA path, called through the C ABI, against the oracle and the golden
vectors of the unmodified reference.  Integer / index results and binary64 scores must be
bit-exact; the fp32 scan scores only have to stay inside the eps band the certificate assumes."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, fusion, index, metrics, retrieval, synth

pytestmark = pytest.mark.gpu

PATHS = [_lib.PATH_WARP, _lib.PATH_MMA]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _lib.load()


def path_ok(path, D, k=10, Q=64, M=1 << 20, G=1, equal=False):
    """The tcgen05 kernel needs cc 10.x and enough gallery tiles for the lists k asks for."""
    if path != _lib.PATH_MMA:
        return True
    info = engine.device_info()
    if not info["has_tcgen05"]:
        return False
    return engine.scan_plan(max(Q, 5), M, D, G, engine.default_k_sel(k), equal)["path"] == _lib.PATH_MMA


def dev(x):
    return engine.quantize(x)


def same_dict(got, want):
    assert set(got) == set(want)
    for k in want:
        assert float(got[k]) == float(want[k]), (k, float(got[k]), float(want[k]))


# --------------------------------------------------------------------------- building blocks
def test_quantize_matches_host_rounding():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 72), dtype=np.float32)
    x[0, 0] = np.float32(1.00390625)      # exact tie between two bf16 values -> even
    got = engine.quantize(x).view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, synth.f32_to_bf16_bits(x))
    n = engine.quantize(x, normalize=True).float().cpu().numpy()
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=4e-3)


def test_score_pairs_bit_exact(small_set):
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    si, st = O.canon_dot64(q, img), O.canon_dot64(q, tgt)
    qq, rr = np.meshgrid(np.arange(q.shape[0]), np.arange(img.shape[0]), indexing="ij")
    pq = torch.from_numpy(qq.ravel().astype(np.int32)).cuda()
    pr = torch.from_numpy(rr.ravel().astype(np.int64)).cuda()
    got = engine.score_pairs(dev(q), dev(img), None, pq, pr).cpu().numpy().reshape(si.shape)
    assert np.array_equal(got, O.canon_fused64(si, None, 1.0))
    got = engine.score_pairs(dev(q), dev(img), dev(tgt), pq, pr, 0.1, 0.9, 0.7).cpu().numpy().reshape(si.shape)
    assert np.array_equal(got, O.canon_fused64(si, st, 0.1, 0.9, 0.7))


@pytest.mark.parametrize("path", PATHS)
def test_scan_scores_stay_inside_eps(path):
    D = 768
    if not path_ok(path, D):
        pytest.skip("tcgen05 path unavailable")
    s = synth.make_retrieval_set(Q=130, M=3000, D=D, seed=5, fused=True, lam=0.3)
    q, img, tgt = dev(s.query), dev(s.image), dev(s.target)
    got = engine.score_matrix(q, img, tgt, 0.1, 0.9, path=path).cpu().numpy().astype(np.float64)
    want = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), 0.1, 0.9)
    err = np.abs(got - want).max()
    print(f"path {path}: max |fp32 scan - canonical| = {err:.3e}")
    assert err < engine.DEFAULT_EPS / 4


# --------------------------------------------------------------------------- top-k
SHAPES = [  # Q, M, D, fused, k
    (1, 1, 8, False, 1), (1, 7, 64, True, 10), (2, 100, 64, False, 10), (3, 1000, 512, True, 10),
    (5, 257, 128, True, 20), (17, 5000, 768, True, 10), (130, 2100, 768, False, 100), (64, 4099, 1024, True, 10),
    (96, 160, 64, True, 5), (300, 700, 192, True, 3),
    # large k on the tensor path: short register lists, many (virtual) parts
    (300, 20000, 128, True, 100), (64, 30000, 64, False, 50), (700, 9000, 256, False, 100),
]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("Q,M,D,fused,k", SHAPES)
def test_scan_topk_matches_canonical(path, Q, M, D, fused, k):
    if not path_ok(path, D, k, Q, M, 2 if fused else 1):
        pytest.skip("tcgen05 path unavailable for this shape")
    s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=100 + Q + M, fused=fused, lam=0.2, diagonal=False)
    w = (0.1, 0.9) if fused else (1.0, 0.0)
    idx, score = engine.scan_topk(dev(s.query), dev(s.image), dev(s.target) if fused else None, w[0], w[1],
                                  k=k, path=path)
    can = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target) if fused else None,
                          w[0], w[1])
    widx, wscore = O.canon_topk(can, k)
    kk = min(k, M)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], widx)
    assert np.array_equal(score.cpu().numpy()[:, :kk], wscore)
    if kk < k:
        assert (idx.cpu().numpy()[:, kk:] == -1).all() and np.isneginf(score.cpu().numpy()[:, kk:]).all()
    assert int((engine.last_flags() != 0).sum()) == 0


@pytest.mark.parametrize("path", PATHS)
def test_scan_topk_exact_ties_lowest_index(path):
    """Duplicate gallery rows give exactly equal scores: the lowest index must win."""
    D = 128
    if not path_ok(path, D):
        pytest.skip("tcgen05 path unavailable")
    g = synth.make_gallery(40, D, 3)
    gal = np.concatenate([g, g, g[:13]], axis=0)          # rows j, j+40 (and j+80) identical
    q = synth.make_queries((g,), np.arange(6), 0.5, 9)
    idx, score = engine.scan_topk(dev(q), dev(gal), k=12, path=path)
    widx, wscore = O.canon_topk(O.canon_fused64(O.canon_dot64(q, gal), None), 12)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(score.cpu().numpy(), wscore)


@pytest.mark.parametrize("path", PATHS)
def test_scan_topk_with_kg_boost(path, small_set):
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    if not path_ok(path, q.shape[1]):
        pytest.skip("tcgen05 path unavailable")
    res, qu, au = small_set["kg_results"], small_set["query_uuids"], small_set["uuids"]
    si, st = O.canon_dot64(q, img), O.canon_dot64(q, tgt)
    r, c, sizes = O.kg_hits_to_pairs(res, qu, au)
    for strategy, params in (("weighted", {"alpha": 0.6, "sparql_weight": 0.4}), ("additive", {"delta": 0.13}),
                             ("adaptive", {"delta": 0.3})):
        alpha, hits = fusion.kg_hits_for_strategy(res, qu, au, strategy, params)
        if strategy == "weighted":
            bonus = O.canon_bonus_matrix(*si.shape, r, c, 0.4, dedupe=True)
        elif strategy == "additive":
            bonus = O.canon_bonus_matrix(*si.shape, r, c, 0.13, dedupe=False)
        else:
            vals = np.array([0.3 * O.omega_for_size(int(sizes[i])) for i in r])
            bonus = O.canon_bonus_matrix(*si.shape, r, c, vals, dedupe=False)
        can = O.canon_fused64(si, st, 0.5, 0.5, alpha, bonus)
        idx, score = engine.scan_topk(dev(q), dev(img), dev(tgt), 0.5, 0.5, alpha, hits, k=10, path=path)
        widx, wscore = O.canon_topk(can, 10)
        assert np.array_equal(idx.cpu().numpy(), widx), strategy
        assert np.array_equal(score.cpu().numpy(), wscore), strategy


def test_certificate_flags_and_retry():
    """k_sel == k leaves no margin: with a huge eps the certificate must refuse, and the retry
    loop must still return the canonical answer."""
    s = synth.make_retrieval_set(Q=9, M=900, D=64, seed=8, fused=False, lam=0.2, diagonal=False)
    q, g = dev(s.query), dev(s.image)
    Q = 9
    score = torch.empty((Q, 4), dtype=torch.float64, device="cuda")
    idx = torch.empty((Q, 4), dtype=torch.int64, device="cuda")
    flags = torch.empty((Q,), dtype=torch.int32, device="cuda")
    ws = engine.workspace_for(Q, 900, 64, 4)
    engine.scan_topk_raw(q, g, None, 1.0, 0.0, 1.0, None, 4, 4, 0.5, 0, score, idx, flags, ws, _lib.PATH_WARP)
    assert (flags.cpu().numpy() & _lib.FLAG_UNCERTIFIED).all()
    i2, s2 = engine.scan_topk(q, g, k=4, k_sel=4, eps=1e-3)
    widx, wscore = O.canon_topk(O.canon_fused64(O.canon_dot64(s.query, s.image), None), 4)
    assert np.array_equal(i2.cpu().numpy(), widx) and np.array_equal(s2.cpu().numpy(), wscore)


# --------------------------------------------------------------------------- ranks and metrics
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(130, 128, 64, False), (333, 2500, 256, True), (257, 300, 128, False)])
def test_rank_targets_padded_query_blocks(path, shape):
    """Query counts that are not multiples of the tensor kernel's block (128 / 256 for CTA pairs):
    padded rows must neither count nor be written (regression: the part-count reduction once wrote
    count[] / flags[] for the padded rows too and clobbered neighbouring allocations)."""
    Q, M, D, fused = shape
    if not path_ok(path, D):
        pytest.skip("tcgen05 path unavailable")
    s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=31, fused=fused, lam=0.15, with_kg=False, diagonal=False)
    q, img = dev(s.query), dev(s.image)
    tgt = dev(s.target) if fused else None
    wa, wb = (0.5, 0.5) if fused else (1.0, 0.0)
    si = O.canon_dot64(s.query, s.image)
    st = O.canon_dot64(s.query, s.target) if fused else None
    tidx = torch.from_numpy(s.target_idx).cuda()
    guard = torch.full((4096,), 7, dtype=torch.int64, device="cuda")      # neighbours of the outputs
    got = engine.rank_targets(q, img, tgt, tidx, wa, wb, path=path).cpu().numpy()
    assert np.array_equal(got, O.canon_rank(O.canon_fused64(si, st, wa, wb), s.target_idx))
    assert bool((guard == 7).all())


@pytest.mark.parametrize("path", PATHS)
def test_rank_targets_matches_canonical(path):
    D = 256
    if not path_ok(path, D):
        pytest.skip("tcgen05 path unavailable")
    s = synth.make_retrieval_set(Q=333, M=2500, D=D, seed=31, fused=True, lam=0.15, with_kg=True, diagonal=False)
    q, img, tgt = dev(s.query), dev(s.image), dev(s.target)
    si, st = O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target)
    tidx = torch.from_numpy(s.target_idx).cuda()
    got = engine.rank_targets(q, img, tgt, tidx, 0.5, 0.5, path=path).cpu().numpy()
    assert np.array_equal(got, O.canon_rank(O.canon_fused64(si, st, 0.5, 0.5), s.target_idx))
    got = engine.rank_targets(q, img, None, tidx, path=path).cpu().numpy()
    assert np.array_equal(got, O.canon_rank(O.canon_fused64(si, None), s.target_idx))
    r, c, _ = O.kg_hits_to_pairs(s.kg_results, s.query_uuids, s.uuids)
    for alpha in (0.9, 0.3):
        a, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted",
                                              {"alpha": alpha, "sparql_weight": 1 - alpha})
        bonus = O.canon_bonus_matrix(*si.shape, r, c, 1 - alpha, dedupe=True)
        want = O.canon_rank(O.canon_fused64(si, st, 0.1, 0.9, a, bonus), s.target_idx)
        got = engine.rank_targets(q, img, tgt, tidx, 0.1, 0.9, a, hits, path=path).cpu().numpy()
        assert np.array_equal(got, want), alpha


def test_metrics_mirror_small_golden(golden, small_set):
    g = golden["metrics_small"]
    q, img, tgt, sim = (small_set[k] for k in ("query", "image", "target", "sim"))
    same_dict(metrics.compute_retrieval_metrics(q, img, prefix="T2I"), g["retrieval_metrics_T2I"])
    same_dict(metrics.compute_retrieval_metrics(q, tgt, k_values=[1, 3, 7]), g["retrieval_metrics_noprefix_k"])
    same_dict(metrics.compute_retrieval_metrics_final(q, tgt, img), g["final_05_05"])
    same_dict(metrics.compute_retrieval_metrics_final(q, tgt, img, prefix="F", t2i_weight=0.1, t2t_weight=0.9),
              g["final_01_09"])
    sq = [small_set[k] for k in ("sq_query", "sq_target", "sq_image")]
    same_dict(metrics.compute_all_retrieval_metrics(*sq), g["all"])
    same_dict(metrics.compute_all_retrieval_metrics(*sq, tasks=["T2I", "T2T"], compute_recall=False),
              g["all_T2I_T2T_mrr_only"])
    same_dict(metrics.compute_training_metrics(*sq), g["training"])
    same_dict(metrics.compute_recall_at_k(sim), g["recall_at_k_matrix"])
    same_dict(metrics.compute_mrr_and_mean_rank(sim), g["mrr_matrix"])
    same_dict(metrics.compute_retrieval_metrics_fusion(sim, prefix="X"), g["metrics_fusion_matrix"])
    same_dict(fusion.evaluate_retrieval(sim), g["evaluate_retrieval"])
    got = metrics.compute_retrieval_metrics(q, img)
    assert all(isinstance(v, np.float64) for v in got.values())
    # the three deprecated shims (metrics.py:285-352): first text variant against itself and against the images
    sq_q, sq_t, sq_i = sq
    same_dict(metrics.compute_metrics_multi_mode(sq_i, [sq_t, sq_q]), g["shim_multi_mode"])
    same_dict(metrics.compute_metrics_single_4train(sq_i, [sq_t]), g["shim_single_4train"])
    same_dict(metrics.compute_metrics_multi_4train(sq_i, [sq_q, sq_t]), g["shim_multi_4train"])


def test_metrics_mirror_mid_golden(golden):
    """1500 x 1500 x 128 set regenerated from its seed; goldens come from the reference."""
    g = golden["metrics_mid"]
    s = synth.make_retrieval_set(Q=1500, M=1500, D=128, seed=23, fused=True, lam=0.4, with_kg=True)
    same_dict(metrics.compute_all_retrieval_metrics(s.query, s.target, s.image), g["all"])
    for wi, wt in ((0.5, 0.5), (0.1, 0.9)):
        same_dict(metrics.compute_retrieval_metrics_final(s.query, s.target, s.image, t2i_weight=wi, t2t_weight=wt),
                  g[f"final_{wi}_{wt}"])
        for alpha in (0.9, 0.5, 0.1):
            got = fusion.evaluate_fused(s.query, s.target, s.image, s.kg_results, s.query_uuids, s.uuids, wi, wt,
                                        "weighted", {"alpha": alpha, "sparql_weight": 1 - alpha})
            same_dict(got, g[f"sweep_{wi}_{wt}_alpha{alpha}"])


def test_dense_fusion_bit_exact(golden, small_set):
    sim = small_set["sim"]
    a = (sim, small_set["kg_results"], small_set["query_uuids"], small_set["uuids"])
    got = {
        "weighted_default": fusion.weighted_fusion(*a),
        "weighted_09_01": fusion.weighted_fusion(*a, alpha=0.9, sparql_weight=1 - 0.9),
        "weighted_renorm": fusion.weighted_fusion(*a, alpha=0.6, sparql_weight=0.6),
        "additive_default": fusion.additive_bonus_fusion(*a),
        "additive_013": fusion.additive_bonus_fusion(*a, delta=0.13),
        "adaptive_default": fusion.adaptive_additive_fusion(*a),
        "adaptive_custom": fusion.adaptive_additive_fusion(*a, delta=0.3, size_thresholds={2: 0.9, 10: 0.4, 25: 0.05}),
        "dispatch_weighted": fusion.fuse_clip_and_text2sparql(*a, fusion_strategy="weighted",
                                                              fusion_params={"alpha": 0.4, "sparql_weight": 0.6}),
        "dispatch_additive": fusion.fuse_clip_and_text2sparql(*a, fusion_strategy="additive"),
        "dispatch_adaptive": fusion.fuse_clip_and_text2sparql(*a, fusion_strategy="adaptive",
                                                              fusion_params={"delta": 0.25}),
    }
    for name, mat in got.items():
        want = small_set["fusion_" + name]
        assert isinstance(mat, np.ndarray) and mat.dtype == np.float32 and mat.shape == want.shape
        assert np.array_equal(mat.view(np.uint32), want.view(np.uint32)), name
        same_dict(fusion.evaluate_retrieval(mat), golden["fusion_small_metrics"][name])
    assert np.array_equal(sim, small_set["sim"])                       # input untouched
    with pytest.raises(ValueError):
        fusion.fuse_clip_and_text2sparql(*a, fusion_strategy="nope")
    with pytest.raises(AssertionError):
        fusion.weighted_fusion(sim, a[1], a[2][:-1], a[3])
    with pytest.raises(AssertionError):
        fusion.additive_bonus_fusion(sim, a[1], a[2], a[3][:-1])


def test_matrix_rank_topk_ties_nan():
    s = np.array([[0.5, 0.9, 0.9, np.nan, 0.1, 0.3], [np.nan, 0.2, 0.2, 0.2, np.nan, -1.0]], np.float32)
    for cols in ([2, 4], [3, 0], [0, 1]):
        got = engine.matrix_rank(s, torch.tensor(cols)).cpu().numpy()
        assert got.tolist() == O.canon_rank(s.astype(np.float64), np.array(cols)).tolist()
    idx, val = engine.matrix_topk(s, 4)
    widx, _ = O.canon_topk(s.astype(np.float64), 4)
    assert np.array_equal(idx.cpu().numpy(), widx)
    rng = np.random.default_rng(1)
    big = rng.standard_normal((33, 5000)).astype(np.float32)
    big[:, 100:200] = big[:, 300:400]                 # exact ties
    idx, val = engine.matrix_topk(big, 20)
    widx, wval = O.canon_topk(big.astype(np.float64), 20)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(val.cpu().numpy(), wval.astype(np.float32))
    t = rng.integers(0, 5000, size=33)
    assert np.array_equal(engine.matrix_rank(big, torch.from_numpy(t)).cpu().numpy(),
                          O.canon_rank(big.astype(np.float64), t))


def test_metrics_reduce_device_equals_host_and_numpy():
    rng = np.random.default_rng(2)
    for n in (1, 9, 128, 129, 1000, 4300, 70001):
        r = rng.integers(1, 50000, size=n).astype(np.int64)
        ks = [1, 5, 10, 20]
        h, s, rr = engine.metrics_reduce(torch.from_numpy(r).cuda(), ks)
        h2, s2, rr2 = engine.metrics_reduce_host(r, ks)
        assert h.tolist() == h2.tolist() and s == s2 and rr == rr2
        assert rr == np.sum(1.0 / r)


@pytest.mark.parametrize("R,Q,k", [(4, 19, 10), (8, 33, 100), (2, 5, 1), (3, 7, 37)])
def test_merge_topk(R, Q, k):
    """Per-shard lists as kemr_scan_topk writes them: ordered by (score desc, idx asc), empty slots (-1) at the end."""
    rng = np.random.default_rng(4 + R)
    sc = np.round(rng.standard_normal((R, Q, k)), 1)               # plenty of exact ties
    ix = rng.permutation(R * Q * k).reshape(R, Q, k).astype(np.int64)
    for r in range(R):
        for qi in range(Q):
            order = sorted(range(k), key=lambda j: (-sc[r, qi, j], ix[r, qi, j]))
            sc[r, qi], ix[r, qi] = sc[r, qi, order], ix[r, qi, order]
    ix[1, :, (7 * k) // 10:] = -1                                  # a short shard
    sc[1, :, (7 * k) // 10:] = -np.inf
    if R > 2:
        ix[2, 0, :] = -1                                           # an empty list
        sc[2, 0, :] = -np.inf
    oi, os_ = engine.merge_topk(torch.from_numpy(sc).cuda(), torch.from_numpy(ix).cuda(), k)
    for qi in range(Q):
        cand = [(-(sc[r, qi, j]), ix[r, qi, j]) for r in range(R) for j in range(k) if ix[r, qi, j] >= 0]
        cand.sort()
        want_i = [c[1] for c in cand[:k]] + [-1] * max(0, k - len(cand))
        want_s = [-c[0] for c in cand[:k]] + [-np.inf] * max(0, k - len(cand))
        assert oi[qi].cpu().tolist() == want_i
        assert os_[qi].cpu().tolist() == want_s


def test_host_index_equals_device_index(small_set):
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    gi = index.GalleryIndex(img, tgt, uuids=small_set["uuids"])
    hi = index.HostIndex(img, tgt, max_queries=128, max_k=20)
    i1, s1 = gi.search(q, k=10, t2i_weight=0.3, t2t_weight=0.7)
    i2, s2, f2 = hi.search(q, k=10, t2i_weight=0.3, t2t_weight=0.7)
    assert np.array_equal(i1.cpu().numpy(), i2) and np.array_equal(s1.cpu().numpy(), s2) and not f2.any()
    lists = [small_set["kg_results"].get(u, []) for u in small_set["query_uuids"]]
    hits = gi.hits_from_uuid_lists(lists, 0.2)
    i3, s3 = gi.search(q, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits=hits)
    csr = (hits.rowptr.cpu().numpy(), hits.col.cpu().numpy(), hits.bonus.cpu().numpy())
    i4, s4, _ = hi.search(q, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits_csr=csr)
    assert np.array_equal(i3.cpu().numpy(), i4) and np.array_equal(s3.cpu().numpy(), s4)
    hi.close()


@pytest.mark.parametrize("locked", [True, False])
def test_host_search_pipelined_chunks_and_small_batches_equal_the_device_search(locked):
    """kemr_index_search_host: pageable batches of 512 queries and more travel as chunks (staging memcpy, PCIe transfer and
    scan of consecutive chunks overlap); one or two queries take the one-copy, one-launch route with their KG hits
    pre-scored by the scanning CTAs.  Every route must return what the device-resident search returns, bit for bit
    (page-locked caller buffers are used in place, pageable ones staged)."""
    s = synth.make_retrieval_set(Q=1100, M=9000, D=256, seed=21, fused=True, lam=0.2, diagonal=False, with_kg=True)
    gi = index.GalleryIndex(s.image, s.target, uuids=s.uuids)
    hi = index.HostIndex(s.image, s.target, max_queries=1100, max_k=10)
    alpha, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted", {"alpha": 0.8, "sparql_weight": 0.2})
    rp, cc, bb = hits.rowptr.cpu().numpy(), hits.col.cpu().numpy(), hits.bonus.cpu().numpy()
    mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()) if locked else (lambda a: np.ascontiguousarray(a))
    for Q in (1, 2, 3, 300, 512, 777, 1100):
        qh = mk(s.query[:Q])
        out = (mk(np.zeros((Q, 10), np.int64)), mk(np.zeros((Q, 10), np.float64)), mk(np.zeros((Q,), np.int32)))
        # plain fused scan
        want_i, want_s = gi.search(s.query[:Q], k=10, t2i_weight=0.4, t2t_weight=0.6)
        got_i, got_s, fl = hi.search(qh, k=10, t2i_weight=0.4, t2t_weight=0.6, out=out)
        assert np.array_equal(want_i.cpu().numpy(), got_i) and np.array_equal(want_s.cpu().numpy(), got_s) and not fl.any(), Q
        # with the KG boost (CSR rows of the first Q queries)
        sub = engine.KGHits(hits.rowptr[:Q + 1].clone(), hits.col[:int(rp[Q])].clone(), hits.bonus[:int(rp[Q])].clone(),
                            hits.max_per_query) if hasattr(engine, "KGHits") else None
        want_i, want_s = gi.search(s.query[:Q], k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=alpha, hits=sub)
        csr = (rp[:Q + 1].copy(), cc[:int(rp[Q])].copy(), bb[:int(rp[Q])].copy())
        got_i, got_s, fl = hi.search(qh, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=alpha, hits_csr=csr, out=out)
        assert np.array_equal(want_i.cpu().numpy(), got_i) and np.array_equal(want_s.cpu().numpy(), got_s) and not fl.any(), Q
        if Q >= 512:                                          # bf16 bit patterns take the same chunked route
            qb = mk(synth.f32_to_bf16_bits(s.query[:Q]))
            got_i, got_s, fl = hi.search(qb, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=alpha, hits_csr=csr, out=out)
            assert np.array_equal(want_i.cpu().numpy(), got_i) and np.array_equal(want_s.cpu().numpy(), got_s), Q
    hi.close()


def test_fused_search_accepts_a_csr_window_of_a_larger_hit_table(small_set):
    """One or two queries take the fused streaming kernel, whose scanning CTAs pre-score the KG hits into a scratch array
    indexed from the CSR's FIRST entry: a window `rowptr[a:b+1]` into a larger table (absolute offsets, as the chunked
    host-buffer search passes it) must give the same result as a re-based CSR of the same queries."""
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    gi = index.GalleryIndex(img, tgt, uuids=small_set["uuids"])
    lists = [small_set["kg_results"].get(u, []) for u in small_set["query_uuids"]]
    hits = gi.hits_from_uuid_lists(lists, 0.2)
    for a0, n in ((7, 1), (20, 2), (len(lists) - 2, 2)):
        window = engine.KGHits(hits.rowptr[a0:a0 + n + 1], hits.col, hits.bonus, hits.max_per_query)
        rebased = hits.subset(torch.arange(a0, a0 + n, device="cuda"))
        assert int(window.rowptr[0]) > 0 and int(rebased.rowptr[0]) == 0
        i1, s1 = gi.search(q[a0:a0 + n], k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits=window)
        i2, s2 = gi.search(q[a0:a0 + n], k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits=rebased)
        assert torch.equal(i1, i2) and torch.equal(s1, s2)
        can = O.canon_fused64(O.canon_dot64(q[a0:a0 + n], img), O.canon_dot64(q[a0:a0 + n], tgt), 0.5, 0.5)
        final = 0.8 * can
        rp, cc, bb = (x.cpu().numpy() for x in (rebased.rowptr, rebased.col, rebased.bonus))
        for i in range(n):
            for h in range(int(rp[i]), int(rp[i + 1])):
                final[i, cc[h]] = final[i, cc[h]] + bb[h]
        widx, wscore = O.canon_topk(final, 10)
        assert np.array_equal(i1.cpu().numpy(), widx) and np.array_equal(s1.cpu().numpy(), wscore)


def test_submit_wait_on_two_lanes_equals_the_blocking_search(small_set):
    """kemr_index_share / kemr_index_submit_host / kemr_index_wait: two lanes over one resident index, searches kept two
    deep; every result equals the blocking call's, for page-locked and pageable buffers, with KG hits, and a second
    submit on a busy lane is refused."""
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    hi = index.HostIndex(img, tgt, max_queries=128, max_k=10)
    lanes = [hi, hi.lane()]
    gi = index.GalleryIndex(img, tgt, uuids=small_set["uuids"])
    lists = [small_set["kg_results"].get(u, []) for u in small_set["query_uuids"]]
    hits = gi.hits_from_uuid_lists(lists, 0.2)
    rp, cc, bb = hits.rowptr.cpu().numpy(), hits.col.cpu().numpy(), hits.bonus.cpu().numpy()
    batches = [(0, 1), (1, 2), (3, 40), (43, 50), (10, 3), (60, 1)]
    for locked in (True, False):
        mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()) if locked else (lambda a: np.ascontiguousarray(a))
        pending = [None, None]
        want_idx, want_score = {}, {}
        for i, (a0, n) in enumerate(batches):
            csr = (rp[a0:a0 + n + 1] - rp[a0], cc[int(rp[a0]):int(rp[a0 + n])].copy(), bb[int(rp[a0]):int(rp[a0 + n])].copy())
            L = lanes[i & 1]
            if pending[i & 1] is not None:
                j, res = pending[i & 1]
                L.wait()
                assert np.array_equal(res[0], want_idx[j]) and np.array_equal(res[1], want_score[j]), (locked, j)
            out = (mk(np.zeros((n, 10), np.int64)), mk(np.zeros((n, 10), np.float64)), mk(np.zeros((n,), np.int32)))
            sub = engine.KGHits(torch.from_numpy(csr[0]).cuda(), torch.from_numpy(csr[1]).cuda(), torch.from_numpy(csr[2]).cuda(), hits.max_per_query)
            wi, ws = gi.search(q[a0:a0 + n], k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits=sub)
            want_idx[i], want_score[i] = wi.cpu().numpy(), ws.cpu().numpy()
            res = L.submit(mk(q[a0:a0 + n]), k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=0.8, hits_csr=csr, out=out)
            pending[i & 1] = (i, res)
        with pytest.raises(_lib.KemrError):
            lanes[0].submit(mk(q[:2]), k=10)                    # lane 0 still has a search in flight
        for lane_no in (0, 1):
            j, res = pending[lane_no]
            lanes[lane_no].wait()
            assert np.array_equal(res[0], want_idx[j]) and np.array_equal(res[1], want_score[j]), (locked, j)
    hi.close()


def test_retrieval_engine_end_to_end(small_set):
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    gi = index.GalleryIndex(img, tgt, uuids=small_set["uuids"])
    table = {f"text {i}": q[i] for i in range(4)}
    clip = retrieval.CLIPRetrieval(retriever=retrieval.CLIPRetriever(gi, encode_text=lambda s: table[s], top_k=25))

    class T2S:
        def retrieval(self, query):
            return [small_set["uuids"][3], "unknown-uuid"]

    eng = retrieval.RetrievalEngine(clip, T2S())
    can = O.canon_fused64(O.canon_dot64(q[:4], img), O.canon_dot64(q[:4], tgt), 0.5, 0.5)
    widx, wscore = O.canon_topk(can, 25)
    for i in range(4):
        plain = eng.retrieve_text_noknowledge(f"text {i}", threshold=-1)
        assert [r["uuid"] for r in plain] == [small_set["uuids"][j] for j in widx[i]]
        assert [r["score"] for r in plain] == wscore[i].tolist()
        fused = eng.retrieve_text(f"text {i}", threshold=-1)
        want = O.ref_fuse_clip_sparql_linear(plain, T2S().retrieval(""), 0.8, 0.2)
        assert fused == want


def test_peer_exchange_two_ranks_in_one_process_equals_single_scan(small_set):
    """The fused result exchange (kemr_peer_*): two ranks emulated in ONE process on one GPU -- each rank's selection
    kernel stores its rows into BOTH exchange buffers, each merge reads its own buffer.  The steps run one after the
    other on one stream, so no kernel ever waits for a kernel that is not already complete.  Must equal the
    single-gallery scan bit for bit, over several epochs (the buffer halves alternate)."""
    import ctypes as C
    lib = _lib.load()
    q, img, tgt = dev(small_set["query"]), dev(small_set["image"]), dev(small_set["target"])
    Q, M, k = q.shape[0], img.shape[0], 10
    cut = 71
    shards = [(img[:cut].contiguous(), tgt[:cut].contiguous(), 0), (img[cut:].contiguous(), tgt[cut:].contiguous(), cut)]
    hs = []
    for r in range(2):
        h = C.c_void_p()
        _lib.check(lib.kemr_peer_create(r, 2, Q, k, C.byref(h), None))
        hs.append(h)
    bases = (C.c_void_p * 2)(*[C.c_void_p(lib.kemr_peer_local_buffer(h)) for h in hs])
    for h in hs:
        _lib.check(lib.kemr_peer_connect_pointers(h, bases))
    st = torch.cuda.current_stream().cuda_stream
    ws = engine.workspace_for(Q, M, q.shape[1], 16)
    for step, (wa, wb) in enumerate(((0.5, 0.5), (0.1, 0.9), (1.0, 0.0))):
        want_i, want_s = engine.scan_topk(q, img, tgt, wa, wb, k=k)
        outs = []
        for r, (a, b, base) in enumerate(shards):
            _lib.check(lib.kemr_peer_begin(hs[r], st))
            sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
            ix = torch.empty((Q, k), dtype=torch.int64, device="cuda")
            fl = torch.empty((Q,), dtype=torch.int32, device="cuda")
            engine.scan_topk_raw(q, a, b, wa, wb, 1.0, None, k, 16, engine.DEFAULT_EPS, base, sc, ix, fl, ws)
        for r in range(2):
            os_ = torch.empty((Q, k), dtype=torch.float64, device="cuda")
            oi = torch.empty((Q, k), dtype=torch.int64, device="cuda")
            _lib.check(lib.kemr_peer_merge(hs[r], Q, k, C.c_void_p(os_.data_ptr()), C.c_void_p(oi.data_ptr()), st))
            outs.append((oi, os_))
        for oi, os_ in outs:
            assert torch.equal(oi, want_i) and torch.equal(os_, want_s), f"step {step}"
    torch.cuda.synchronize()
    for h in hs:
        lib.kemr_peer_destroy(h)


def test_prepared_sharded_search_on_one_rank(small_set):
    """distributed.SearchPlan with a world of one (peer exchange with itself / plain copy), eager and as a CUDA graph."""
    from knowledge_enhanced_multimodal_retrieval_b200.distributed import CudaLocal, ShardedGallery
    q = dev(small_set["query"])
    sg = ShardedGallery(CudaLocal(small_set["image"], small_set["target"]), small_set["image"].shape[0])
    want_i, want_s = engine.scan_topk(q, sg.local.image, sg.local.target, 0.3, 0.7, k=10)
    for exchange in ("peer", "nccl"):
        for graph in (False, True):
            plan = sg.plan(Q=q.shape[0], k=10, w_a=0.3, w_b=0.7, exchange=exchange, graph=graph)
            for _ in range(3):
                pi, ps = plan.run(q)
                assert torch.equal(pi, want_i) and torch.equal(ps, want_s) and plan.uncertified() == 0
            plan.close()


def test_degenerate_gallery_falls_back_to_an_exact_ranking():
    """Hundreds of identical rows around the k-th score cannot be certified by any finite selection: the
    search then decides those queries exactly instead of raising -- lowest index first among equal scores, like a
    stable argsort of the reference's matrix."""
    rng = np.random.default_rng(2)
    base = synth.make_gallery(8, 64, 3)
    img = np.repeat(base, 400, axis=0)                                 # 3200 rows, 8 distinct: every score is tied 400 times
    q = synth.round_to_bf16(synth.l2_normalize(base[:6] + 0.1 * rng.standard_normal((6, 64), dtype=np.float32)))
    idx, sc = engine.scan_topk(dev(q), dev(img), k=10)
    can = O.canon_dot64(q, img)
    widx, wsc = O.canon_topk(can, 10)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(sc.cpu().numpy(), wsc)


@pytest.mark.parametrize("path", PATHS)
def test_eps_bound_holds_on_adversarial_inputs(path):
    """The certificate rests on |fp32 scan score - canonical| <= eps (engine.DEFAULT_EPS for unit-norm inputs, derived
    in DESIGN.md §2 from one rounding per accumulator update).  Worst cases for fp32 accumulation: same-sign vectors
    (no cancellation, the running sum grows monotonically), the largest supported D, two galleries sharing one
    accumulator, near-duplicate rows."""
    D, M, Q = 1024, 4096, 130
    if not path_ok(path, D, Q=Q, M=M, G=2, equal=True):
        pytest.skip("tcgen05 path unavailable")
    rng = np.random.default_rng(5)
    pos = np.abs(rng.standard_normal((M, D), dtype=np.float32)) + 0.5            # all components positive
    img = synth.round_to_bf16(synth.l2_normalize(pos))
    tgt = synth.round_to_bf16(synth.l2_normalize(pos[::-1] * 0.9 + 0.1))
    img[1::64] = img[0::64]                                                      # exact duplicates
    nxt = img[2::64].copy()
    img[3::64] = synth.round_to_bf16(nxt + np.float32(2.0 ** -9) * (nxt > 0.03))  # near duplicates: a few bf16 ulps away
    q = synth.round_to_bf16(synth.l2_normalize(np.abs(rng.standard_normal((Q, D), dtype=np.float32)) + 0.5))
    qd, a, b = dev(q), dev(img), dev(tgt)
    for wa, wb in ((0.5, 0.5), (0.1, 0.9), (1.0, 0.0)):
        two = wb != 0.0
        can = O.canon_fused64(O.canon_dot64(q, img), O.canon_dot64(q, tgt) if two else None, wa, wb)
        dense = engine.score_matrix(qd, a, b if two else None, wa, wb, path=path).cpu().numpy()
        err = float(np.abs(dense.astype(np.float64) - can).max())
        assert err < engine.DEFAULT_EPS, (wa, wb, err)
        idx, sc = engine.scan_topk(qd, a, b if two else None, wa, wb, k=10, path=path)
        widx, wsc = O.canon_topk(can, 10)
        assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(sc.cpu().numpy(), wsc)
        print(f"path {path} weights {(wa, wb)}: max |fp32 - canonical| = {err:.3e} (eps {engine.DEFAULT_EPS:.1e})")


@pytest.mark.parametrize("path", PATHS)
def test_unnormalised_embeddings_scale_the_margin(path):
    """Rows of norm ~10 and weights above one: the margin follows eps_for = DEFAULT_EPS * max||q|| *
    (|w_a| max||g_a|| + |w_b| max||g_b||); top-k and ranks stay exact against the oracle, and the fp32 scan scores stay
    inside that margin."""
    D, M, Q = 256, 3000, 150
    if not path_ok(path, D, Q=Q, M=M, G=2):
        pytest.skip("tcgen05 path unavailable")
    rng = np.random.default_rng(9)
    img = synth.round_to_bf16(rng.standard_normal((M, D), dtype=np.float32) * np.float32(10.0 / np.sqrt(D)))
    tgt = synth.round_to_bf16(rng.standard_normal((M, D), dtype=np.float32) * np.float32(7.0 / np.sqrt(D)))
    q = synth.round_to_bf16(img[:Q] * np.float32(0.3) + rng.standard_normal((Q, D), dtype=np.float32) * np.float32(3.0 / np.sqrt(D)))
    qd, a, b = dev(q), dev(img), dev(tgt)
    wa, wb = 1.5, 2.25
    eps = engine.eps_for(qd, a, b, wa, wb)
    nq, na, nb = (float(np.linalg.norm(x.astype(np.float64), axis=1).max()) for x in (q, img, tgt))
    assert eps >= engine.DEFAULT_EPS * nq * (wa * na + wb * nb) and eps < 1.02 * engine.DEFAULT_EPS * nq * (wa * na + wb * nb)
    can = O.canon_fused64(O.canon_dot64(q, img), O.canon_dot64(q, tgt), wa, wb)
    dense = engine.score_matrix(qd, a, b, wa, wb, path=path).cpu().numpy()
    assert float(np.abs(dense.astype(np.float64) - can).max()) < eps
    idx, sc = engine.scan_topk(qd, a, b, wa, wb, k=10, path=path)                # eps=None: scaled automatically
    widx, wsc = O.canon_topk(can, 10)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(sc.cpu().numpy(), wsc)
    assert int((engine.last_flags() != 0).sum()) == 0
    ranks = engine.rank_targets(qd, a, b, torch.arange(Q, device="cuda"), wa, wb, path=path)
    assert np.array_equal(ranks.cpu().numpy(), O.canon_rank(can, np.arange(Q)))


@pytest.mark.parametrize("Q", [2, 150])
def test_host_index_margin_follows_gallery_norms_and_weights(Q):
    """kemr_index_search_host scales its selection margin with the resident galleries' largest row norms (measured at
    kemr_index_create) and the weights, like engine.eps_for: galleries of norm ~10 / ~7 and weights above one, unit
    queries -- the small-batch and the large-batch route both return the oracle's top-k with every query certified."""
    D, M = 256, 3000
    rng = np.random.default_rng(10)
    img = synth.round_to_bf16(rng.standard_normal((M, D), dtype=np.float32) * np.float32(10.0 / np.sqrt(D)))
    tgt = synth.round_to_bf16(rng.standard_normal((M, D), dtype=np.float32) * np.float32(7.0 / np.sqrt(D)))
    q = img[:Q] * np.float32(0.3) + rng.standard_normal((Q, D), dtype=np.float32) * np.float32(3.0 / np.sqrt(D))
    q = synth.round_to_bf16(q / np.linalg.norm(q.astype(np.float64), axis=1, keepdims=True).astype(np.float32) * np.float32(0.995))
    assert float(np.linalg.norm(q.astype(np.float64), axis=1).max()) <= 1.0
    wa, wb = 1.5, 2.25
    hi = index.HostIndex(img, tgt, max_queries=256, max_k=10)
    idx, sc, fl = hi.search(q, k=10, t2i_weight=wa, t2t_weight=wb)
    hi.close()
    can = O.canon_fused64(O.canon_dot64(q, img), O.canon_dot64(q, tgt), wa, wb)
    widx, wsc = O.canon_topk(can, 10)
    assert np.array_equal(idx, widx) and np.array_equal(sc, wsc) and not fl.any()


def test_sharded_index_maps_global_ids_to_its_own_uuid_slice(small_set):
    """A shard with idx_base > 0 returns GLOBAL row ids; CLIPRetriever.search must look them up in the shard's own
    uuid slice."""
    q, img, tgt = small_set["query"], small_set["image"], small_set["target"]
    lo = 57
    gi = index.GalleryIndex(img[lo:], tgt[lo:], uuids=small_set["uuids"][lo:], idx_base=lo)
    r = retrieval.CLIPRetriever(gi, top_k=7)
    can = O.canon_fused64(O.canon_dot64(q[:3], img[lo:]), O.canon_dot64(q[:3], tgt[lo:]), 0.5, 0.5)
    widx, wscore = O.canon_topk(can, 7)
    for i in range(3):
        got = r.search(q[i], alpha=0.5)
        assert [g["uuid"] for g in got] == [small_set["uuids"][lo + j] for j in widx[i]]
        assert [g["score"] for g in got] == wscore[i].tolist()
    with pytest.raises(ValueError):
        r.search("a string query without an encoder")


def test_more_queries_than_candidates_follows_the_reference(golden, small_set):
    """Tall inputs (N > M): rows i >= M have no target column; the reference scores them as 'no recall hit,
    position 1' (metrics.py:37,41,68).  Both the matrix-taking and the embedding-taking mirrors reproduce its dicts."""
    q, img = small_set["query"], small_set["image"][:40]
    tall = (q @ img.T).astype(np.float32)
    same_dict(metrics.compute_retrieval_metrics_fusion(tall), golden["tall_matrix"]["metrics"])
    same_dict(metrics.compute_retrieval_metrics(q, img), golden["tall_matrix"]["embeddings"])
    rec = metrics.compute_recall_at_k(tall)
    mrr = metrics.compute_mrr_and_mean_rank(tall)
    same_dict({**rec, **mrr}, golden["tall_matrix"]["metrics"])
    # rows outside the shard score NaN instead of reading out of bounds
    s = engine.score_pairs(dev(q), dev(img), None, torch.tensor([0, 1], dtype=torch.int32).cuda(),
                           torch.tensor([40, -1]).cuda())
    assert bool(torch.isnan(s).all())


def test_generic_fp32_embeddings_stay_within_the_stated_tolerance(golden):
    """Embeddings that are NOT bf16-representable (what an encoder produces): the engine stores them as bf16
    (round-to-nearest-even) while the reference scores the fp32 values.  Against the UNMODIFIED reference's outputs
    (tests/golden/make_golden.py, 'generic_fp32'): scores of the reference's own top-5 rows within 1e-3 absolute
    (north_star), Recall@K within 2 queries of 200, MRR within 1 point, Mean_Rank within 2 %."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generic_fp32.npz"))
    q, img, tgt = z["query"], z["image"], z["target"]
    G = golden["generic_fp32"]
    assert not np.array_equal(synth.round_to_bf16(q), q)                      # the inputs really are generic fp32
    top5 = np.array(G["top5"])
    n = len(q)
    pq = torch.arange(n, dtype=torch.int32).repeat_interleave(5).cuda()
    got = engine.score_pairs(dev(q), dev(img), dev(tgt), pq, torch.from_numpy(top5.ravel()).cuda(), 0.5, 0.5)
    err = np.abs(got.cpu().numpy().reshape(n, 5) - np.array(G["top5_scores"]))
    assert err.max() < 1e-3, err.max()
    for name, ours in (("final_0.5_0.5", metrics.compute_retrieval_metrics_final(q, tgt, img)),
                       ("T2I", metrics.compute_retrieval_metrics(q, img))):
        ref = G[name]
        for k in ("R@1", "R@5", "R@10", "R@20"):
            assert abs(float(ours[k]) - ref[k]) <= 2 * 100.0 / n + 1e-9, (name, k, float(ours[k]), ref[k])
        assert abs(float(ours["MRR"]) - ref["MRR"]) <= 1.0, (name, float(ours["MRR"]), ref["MRR"])
        assert abs(float(ours["Mean_Rank"]) - ref["Mean_Rank"]) <= 0.02 * ref["Mean_Rank"] + 0.05
    print("generic fp32: max |score - reference| =", err.max())


# --------------------------------------------------------------------------- per-query gated fusion (§8f)
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("Q,M,D,k", [(130, 2000, 128, 10), (300, 3000, 768, 20), (7, 900, 64, 5)])
def test_gated_per_query_weights_match_canonical(path, Q, M, D, k):
    if not path_ok(path, D, k, Q, M, 2):
        pytest.skip("tcgen05 path unavailable for this shape")
    s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=500 + Q, fused=True, lam=0.2, diagonal=True)
    rng = np.random.default_rng(Q)
    gate = rng.uniform(0.02, 0.98, Q).astype(np.float32)
    wa, wb = gate.astype(np.float64), (np.float32(1.0) - gate).astype(np.float64)
    q, img, tgt = dev(s.query), dev(s.image), dev(s.target)
    idx, score = engine.scan_topk(q, img, tgt, wa, wb, k=k, path=path)
    can = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), wa, wb)
    widx, wscore = O.canon_topk(can, k)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(score.cpu().numpy(), wscore)
    assert int((engine.last_flags() != 0).sum()) == 0
    ranks = engine.rank_targets(q, img, tgt, torch.from_numpy(s.target_idx).cuda(), wa, wb, path=path)
    assert np.array_equal(ranks.cpu().numpy(), O.canon_rank(can, s.target_idx))


def test_gated_heads_follow_the_reference_formula():
    from knowledge_enhanced_multimodal_retrieval_b200 import fusion_heads as FH
    s = synth.make_retrieval_set(Q=200, M=200, D=256, seed=77, fused=True, lam=0.3, diagonal=True)
    rng = np.random.default_rng(5)
    w = rng.normal(0, 0.5, 256).astype(np.float32)
    head = FH.SimpleGatedFusion(w, bias=0.25, embed_dim=256)
    q = dev(s.query)
    wa, wb = head.gate_weights(q)
    gate_ref = O.ref_gate_linear(s.query, w, 0.25)
    assert np.abs(wa.cpu().numpy() - gate_ref).max() < 2e-6                     # fp32 sum order differs, not more
    assert np.array_equal(wb.cpu().numpy(), (np.float32(1) - wa.cpu().numpy().astype(np.float32)).astype(np.float64))
    # metrics of the head == metrics of the reference's (N, N) fp32 score matrix (fusion_model.py:16-22 + metrics)
    got = head.evaluate(s.query, s.image, s.target)
    ref = O.ref_gated_scores(s.query, s.image, s.target, wa.cpu().numpy().astype(np.float32))
    want = O.ref_metrics_from_matrix(ref)
    can = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target),
                          wa.cpu().numpy(), wb.cpu().numpy())
    same_dict(got, O.metrics_from_ranks(O.canon_rank(can, s.target_idx)))
    assert np.abs(ref - can).max() < 1e-5                                        # canonical == reference formula
    for key in want:                                                             # at most one near-tie may flip
        assert abs(float(got[key]) - float(want[key])) <= 0.51, (key, got[key], want[key])
    # defaults of the two simple heads: gate = sigmoid(sum(q)) and sigmoid(-2)
    g0, _ = FH.SimpleGatedFusionWithBias(embed_dim=256).gate_weights(q)
    assert np.allclose(g0.cpu().numpy(), 1.0 / (1.0 + np.exp(2.0)), atol=1e-7)
    # MLP gate and bilinear head run end to end and agree with a direct evaluation of their formulas
    mlp = FH.GatedFusionHead(rng.normal(0, 0.1, (128, 256)), rng.normal(0, 0.1, 128), rng.normal(0, 0.1, (1, 128)), 0.1)
    idx, sc = mlp.search(s.query, s.image, s.target, k=5)
    ga, gb = mlp.gate_weights(q)
    can2 = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), ga.cpu().numpy(), gb.cpu().numpy())
    widx, wsc = O.canon_topk(can2, 5)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(sc.cpu().numpy(), wsc)
    bil = FH.BilinearFusionHead(np.eye(256, dtype=np.float32), np.eye(256, dtype=np.float32), alpha=0.0)
    bil.project(s.image, s.target)
    same_dict(bil.evaluate(s.query), metrics.compute_retrieval_metrics_final(s.query, s.target, s.image))


# --------------------------------------------------------------------------- grouped ground truth (§8f)
def test_grouped_ground_truth_metrics():
    """N queries against an N x 4 candidate pool, candidate c belongs to artefact c // 4
    (baselines/evaluate_text_models.py:176-224)."""
    N, D = 300, 256
    base = synth.make_gallery(N, D, 31)
    rng = np.random.default_rng(32)
    cands = synth.round_to_bf16(synth.l2_normalize(np.repeat(base, 4, axis=0) * 0.6
                                                   + rng.normal(0, 1.0 / np.sqrt(D), (4 * N, D)).astype(np.float32)))
    query = synth.round_to_bf16(synth.l2_normalize(base * 0.5 + rng.normal(0, 1.0 / np.sqrt(D), (N, D)).astype(np.float32)))
    t2a = np.repeat(np.arange(N), 4)
    perm = rng.permutation(4 * N)                                  # positives need not be contiguous
    cands, t2a = cands[perm], t2a[perm]
    q, c = dev(query), dev(cands)
    ranks = metrics.grouped_ranks(q, c, t2a)
    can = O.canon_dot64(query, cands)
    assert np.array_equal(ranks.cpu().numpy(), O.canon_grouped_rank(can, t2a))
    got = metrics.compute_grouped_retrieval_metrics(query, cands, t2a)
    same_dict(got, O.metrics_from_ranks(O.canon_grouped_rank(can, t2a), prefix="T2T"))
    want = O.ref_grouped_metrics(query, cands, t2a)                # the reference's loops (fp32, unstable sort)
    for key in want:
        assert abs(float(got[key]) - float(want[key])) <= 0.34, (key, got[key], want[key])
    # multi mode: two "variants" of queries against the same pool, reduced over the concatenated ranks
    q2 = np.concatenate([query, query[::-1]])
    qa = np.concatenate([np.arange(N), np.arange(N)[::-1]])
    r2 = metrics.grouped_ranks(dev(q2), c, t2a, qa)
    assert np.array_equal(r2.cpu().numpy()[:N], ranks.cpu().numpy()) and np.array_equal(r2.cpu().numpy()[N:], ranks.cpu().numpy()[::-1])
