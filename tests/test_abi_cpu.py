"""CPU: the C-ABI library builds/loads and exports every function include/tempme_b200.h declares
(no compute calls: there is no GPU here), plus host-side argument validation that needs no device."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tempme_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from tempme_b200 import _lib
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tempme_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), "ctypes SIGNATURES out of sync with the header"
    assert L.tm_version() == 100


def test_argument_validation_without_device():
    from tempme_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    bad = np.array([5], np.int32); z = np.zeros(1, np.int32); t = np.zeros(1, np.float64)
    rc = L.tm_graph_create(2, 1, _lib.ptr(bad), _lib.ptr(z), _lib.ptr(z), _lib.ptr(t), 0, C.byref(h))
    assert rc == -4 and b"outside" in L.tm_last_error()          # TM_ERR_NODE_RANGE
    rc = L.tm_graph_create(2, -1, None, None, None, None, 0, C.byref(h))
    assert rc == -1
    assert L.tm_sample_walks(None, 1, 1, 1, None, None, None, None, 0, 0, None, None, None, None, None, None, None, None, None, None, None) == -1


def test_encoder_blob_layout_and_pack():
    from tempme_b200 import _lib
    L = _lib.lib()
    d = _lib.EncoderDesc(172, 1, 64, 1, 1)
    n = L.tm_encoder_blob_floats(C.byref(d))
    # Wt[Kp][Np] + bias[Np] per Linear, K padded to 4 and N to 32
    ev = 172 + 3 + 1
    exp = (176 * 192 + 192) + (172 * 64 + 64) + (64 * 64 + 64) + 2 * (128 * 128 + 128) + (128 * 64 + 64) + (64 * 64 + 64) \
        + (76 * 96 + 96) + (76 * 64 + 64) + (64 * 32 + 32) + 2 * 192
    assert ev == 176 and n > exp and n % 4 == 0     # fp32 layout first, then the pre-split tcgen05 operand tiles
    rng = np.random.default_rng(0)
    shapes = [(172, 176), (172,), (64, 172), (64,), (64, 64), (64,), (128, 128), (128,), (128, 128), (128,), (64, 128), (64,),
              (64, 64), (64,), (76, 76), (76,), (64, 76), (64,), (1, 64), (1,), (172,), (172,)]
    arrs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    prm = _lib.EncoderParams(*[_lib.ptr(a) for a in arrs])
    blob = np.full(n, np.nan, np.float32)
    assert L.tm_encoder_pack(C.byref(d), C.byref(prm), _lib.ptr(blob)) == 0
    assert np.isfinite(blob).all()
    tcb = blob[exp:]
    assert (np.frombuffer(tcb[:176 * 32].tobytes(), np.uint32) & 0x1FFF == 0).all()     # first hi tile: tf32-exact values
    wt = blob[:176 * 192].reshape(176, 192)
    assert np.array_equal(wt[:, :172], arrs[0].T) and (wt[:, 172:] == 0).all()
    assert np.array_equal(blob[176 * 192:176 * 192 + 172], arrs[1])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tempme_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), f
                assert "tempme_oracle" not in s, f
