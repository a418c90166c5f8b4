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


def test_new_entry_points_validate_arguments_without_a_gpu():
    import ctypes as C
    lib = _lib.load()
    null = None
    one = (C.c_int64 * 2)(0, 0)
    # hits_build_csr: null pointers, negative sizes, inverted row range
    assert lib.kemr_hits_build_csr(null, null, null, 4, 0, 10, 0, null, null, null, null, null, 0, null) == 1
    assert lib.kemr_hits_build_csr(one, one, one, 0, 0, 10, 0, one, null, null, one, one, 4096, null) == 1
    assert lib.kemr_hits_build_csr(one, one, one, 1, 10, 5, 0, one, null, null, one, one, 4096, null) == 1
    assert b"hits_build_csr" in lib.kemr_last_error()
    assert lib.kemr_hits_workspace_bytes(1000) >= 8000
    # idmap: bad offsets are refused, empty maps work
    h = C.c_void_p()
    bad = (C.c_int64 * 3)(0, 5, 2)
    assert lib.kemr_idmap_create(b"abcde", bad, 2, C.byref(h)) == 1
    zero = (C.c_int64 * 1)(0)
    assert lib.kemr_idmap_create(b"", zero, 0, C.byref(h)) == 0
    out = (C.c_int64 * 1)(7)
    off = (C.c_int64 * 2)(0, 3)
    assert lib.kemr_idmap_lookup(h, b"abc", off, 1, 1, out) == 0 and out[0] == -1
    lib.kemr_idmap_destroy(h)
    # gated calls without weight arrays, store calls with bad shapes
    assert lib.kemr_scan_topk_gated(null, 1, null, null, 1, 8, null, null, 1.0, null, null, null, 0, 1, 1, 1e-5, 0,
                                    null, null, null, null, null, 0, 0, null) == 1
    assert lib.kemr_gate_linear(null, 1, 8, null, 0.0, null, null, null) == 1
    assert lib.kemr_store_write(b"/tmp/x.kemr", null, 3, 8) == 1
    assert lib.kemr_store_write(b"/nonexistent-dir/x.kemr", (C.c_uint16 * 8)(), 1, 8) == 1
    assert lib.kemr_store_load(b"/tmp/definitely-missing.kemr", 0, 1, (C.c_uint16 * 8)(), null) == 1
