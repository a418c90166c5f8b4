"""Host side of the data path around the scan (no GPU): uuid map and the embedding-store file format."""
import os

import numpy as np
import pytest

from knowledge_enhanced_multimodal_retrieval_b200 import _lib, store, synth


def test_idmap_matches_python_dict_semantics():
    uuids = [f"u{i}" for i in range(1000)] + ["dup", "x/y", "dup", ""]
    m = store.IdMap(uuids)
    ref = {u: j for j, u in enumerate(uuids)}                         # fusion.py:62: the last duplicate wins
    keys = ["u0", "u999", "http://x/u17", "a/b/c/u5", "nope", "dup", "http://host/dup", "x/y", "", "trailing/"]
    got = m.rows(keys, normalize_uri=True)
    want = [ref.get(k.split("/")[-1], -1) for k in keys]              # fusion.py:76-78
    assert got.tolist() == want
    raw = m.rows(["x/y", "http://x/u17", "u17"], normalize_uri=False)
    assert raw.tolist() == [ref["x/y"], -1, 17]
    assert m.rows([]).tolist() == []
    m.close()


def test_read_text2sparql_results_layout(tmp_path):
    d = tmp_path / "results"
    d.mkdir()
    (d / "q1.txt").write_text("http://x/u1\nu2 \n\n")
    (d / "q2.v2.txt").write_text("")
    got = store.read_text2sparql_results(str(d))
    assert got == {"q1": ["http://x/u1", "u2", ""], "q2": []}         # evaluator.py:46-50: strip only, keep blanks


def test_store_file_format_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    bits = synth.f32_to_bf16_bits(synth.round_to_bf16(rng.normal(size=(37, 64)).astype(np.float32)))
    lib = _lib.load()
    import ctypes as C
    path = str(tmp_path / "image.kemr").encode()
    _lib.check(lib.kemr_store_write(path, bits.ctypes.data_as(C.c_void_p), 37, 64))
    rows, dim = C.c_int64(), C.c_int()
    _lib.check(lib.kemr_store_info(path, C.byref(rows), C.byref(dim)))
    assert (rows.value, dim.value) == (37, 64)
    raw = open(path, "rb").read()
    assert raw[:8] == b"KEMRSTOR" and len(raw) == 64 + 37 * 64 * 2
    assert np.array_equal(np.frombuffer(raw[64:], dtype=np.uint16).reshape(37, 64), bits)
    with open(path, "r+b") as f:                                         # truncated file is refused
        f.truncate(64 + 10)
    with pytest.raises(_lib.KemrError):
        _lib.check(lib.kemr_store_info(path, C.byref(rows), C.byref(dim)))
    with pytest.raises(_lib.KemrError):
        _lib.check(lib.kemr_store_info(str(tmp_path / "missing.kemr").encode(), C.byref(rows), C.byref(dim)))
