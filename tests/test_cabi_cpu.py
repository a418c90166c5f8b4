"""CPU: the C-ABI library loads, exports every symbol include/kemr.h declares, its host-side
entry points work without a GPU, and the host logic of the Python mirror matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, fusion, retrieval, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    hdr = open(os.path.join(ROOT, "include", "kemr.h")).read()
    declared = set(re.findall(r"\b(kemr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"kemr_index"}            # struct tag, not a function
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().kemr_abi_version() == 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.KemrError):
        engine.quantize(np.zeros((2, 8), np.float32))
    with pytest.raises(_lib.KemrError):
        from knowledge_enhanced_multimodal_retrieval_b200 import metrics
        metrics.compute_retrieval_metrics(np.zeros((2, 8), np.float32), np.zeros((2, 8), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "knowledge_enhanced_multimodal_retrieval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/_ref" not in src and "/root/reference" not in src, f


def test_metrics_reduce_host_matches_numpy():
    rng = np.random.default_rng(3)
    for n in [1, 2, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 4300, 8193, 100003]:
        r = rng.integers(1, 60000, size=n).astype(np.int64)
        ks = [1, 5, 10, 20]
        hits, s, rr = engine.metrics_reduce_host(r, ks)
        assert [int(h) for h in hits] == [int((r <= k).sum()) for k in ks]
        assert s == float(r.sum())
        assert rr / np.float64(n) * 100.0 == np.mean(1.0 / r) * 100.0       # metrics.py:70
        assert s / np.float64(n) == np.mean(r)                               # metrics.py:71


def test_reduce_host_rank_zero_is_the_reference_no_match_case(golden, small_set):
    """More queries than candidates: the reference has no column i for rows i >= M -- never a recall hit, position
    argmax(all False) + 1 = 1 (metrics.py:41,68).  The reductions read rank 0 as exactly that: the host twin applied
    to canonical ranks of the first M rows + zeros reproduces the unmodified reference's dict on the tall matrix."""
    from oracle import oracle as O
    q, img = small_set["query"], small_set["image"][:40]
    can = O.canon_dot64(q[:40], img)
    ranks = np.concatenate([O.canon_rank(can, np.arange(40)), np.zeros(len(q) - 40, np.int64)])
    ks = [1, 5, 10, 20]
    hits, s, rr = engine.metrics_reduce_host(ranks, ks)
    n = np.float64(len(q))
    got = {f"R@{k}": float(np.float64(h) / n * 100.0) for k, h in zip(ks, hits)}
    got["MRR"] = float(np.float64(rr) / n * 100.0)
    got["Mean_Rank"] = float(np.float64(s) / n)
    assert got == golden["tall_matrix"]["metrics"] == golden["tall_matrix"]["embeddings"]


def test_engine_list_fusion_matches_reference(golden):
    eng = retrieval.RetrievalEngine()
    for c in golden["engine"]["fuse_cases"]:
        assert eng._fuse_clip_sparql_linear(c["clip"], c["sparql"], c["alpha"], c["beta"]) == c["out"]
    ci, calls = golden["engine"]["call_inputs"], golden["engine"]["calls"]

    class Fixed:
        def __init__(self, res):
            self.res = res

        def retrieval(self, query, alpha=0.5):
            return self.res

    class FixedT2S(Fixed):
        def retrieval(self, query):
            return self.res

    eng = retrieval.RetrievalEngine(Fixed(ci["clip"]), FixedT2S(ci["sparql"]))
    assert eng.retrieve_text("q") == calls["retrieve_text_default"]
    assert eng.retrieve_text("q", alpha=0.6, beta=0.4, threshold=0.3) == calls["retrieve_text_thr"]
    assert eng.retrieve_text_noknowledge("q") == calls["noknowledge_default"]
    assert eng.retrieve_text_noknowledge("q", threshold=0.25) == calls["noknowledge_thr"]
    assert eng._fuse_clip_sparql_linear([], ["x"]) == []


def test_kg_host_logic_matches_oracle(small_set):
    res, qu, au = small_set["kg_results"], small_set["query_uuids"], small_set["uuids"]
    cols, sizes = engine.kg_pairs(res, qu, au)
    r, c, sz = O.kg_hits_to_pairs(res, qu, au)
    assert [j for cs in cols for j in cs] == c.tolist()
    assert [i for i, cs in enumerate(cols) for _ in cs] == r.tolist()
    assert sizes == sz.tolist()
    assert engine.uri_tail("http://x/y/u1") == "u1" and engine.uri_tail("u1") == "u1"
    assert fusion._omega(1, fusion._DEFAULT_OMEGA) == 1.0 and fusion._omega(51, fusion._DEFAULT_OMEGA) == 0.1
    assert fusion._omega(30, {2: 0.9, 10: 0.4, 25: 0.05}) == 0.0 == O.omega_for_size(30, {2: 0.9, 10: 0.4, 25: 0.05})


def test_order_key_roundtrip():
    # mirrors common.cuh::order_f32 so the host can reason about candidate keys
    def order(f):
        u = np.float32(f).view(np.uint32)
        return (~u) & np.uint32(0xffffffff) if u & np.uint32(0x80000000) else u | np.uint32(0x80000000)
    xs = np.array([-np.inf, -1.0, -1e-30, -0.0, 0.0, 1e-30, 0.5, 1.0, np.inf], np.float32)
    keys = [int(order(x)) for x in xs]
    assert keys == sorted(keys)


def test_integration_md_ctypes_stub_matches_the_abi():
    """The stub INTEGRATION.md shows a maintainer (section 2) must keep working against the header: same entry points,
    same argument order.  Without a GPU the calls fail in CUDA, not in argument marshalling."""
    import ctypes as C
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    used = set(re.findall(r"lib\.(kemr_\w+)", text))
    assert used and used <= set(_lib.EXPORTS), used - set(_lib.EXPORTS)
    header = open(os.path.join(ROOT, "include", "kemr.h")).read()
    for name in re.findall(r"`(kemr_\w+)`", text):
        base = name.rstrip("*")
        assert any(e.startswith(base.replace("_*", "")) for e in _lib.EXPORTS) or base in header, name
    lib = C.CDLL(_lib.LIB_PATH)
    lib.kemr_last_error.restype = C.c_char_p
    img = np.zeros((4, 8), np.uint16)
    h = C.c_void_p()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.kemr_index_create(vp(img), vp(img), C.c_int64(4), 8, 16, 10, C.byref(h))      # the stub's call, verbatim shape
    if rc != 0:
        assert rc == 2 and b"cuda" in lib.kemr_last_error().lower()                         # KEMR_ERR_CUDA on a CPU-only box
