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


def test_detect_fanout_host_logic():
    """TempME.detect_fanout (the walk-layout hint of tm_encoder_desc.walk_fanout) on CPU tensors: find_k_walks' layout w = i1 * N2 + j gives N2;
    shuffled walks give 1; small calls are skipped."""
    import torch
    from tempme_b200.explainer import TempME
    B, n = 4, 5
    for N2 in (1, 2, 3, 5):
        W = n * N2
        first = torch.arange(B * n, dtype=torch.int32).view(B, n)              # one distinct first-hop event per slot
        e = torch.zeros((B, W, 3), dtype=torch.int32); nodes = torch.zeros((B, W, 6), dtype=torch.int32)
        e[:, :, 2] = first.repeat_interleave(N2, dim=1)
        nodes[:, :, 4] = 7; nodes[:, :, 5] = first.repeat_interleave(N2, dim=1) + 100
        e[:, :, 1] = torch.arange(W, dtype=torch.int32)                         # the other positions differ per walk
        TempME._fanout_seen.clear()
        assert TempME.detect_fanout(e, nodes) == N2
        assert TempME.detect_fanout(e, nodes) == N2                             # cached candidate first
        assert TempME.detect_fanout(e, nodes, min_motifs=10 ** 6) == 1          # below the threshold: no test
        if N2 > 1:
            perm = torch.randperm(W, generator=torch.Generator().manual_seed(N2))
            if not torch.equal(e[:, perm, 2].view(B, W // N2, N2), e[:, perm, 2].view(B, W // N2, N2)[:, :, :1].expand(B, W // N2, N2)):
                assert TempME.detect_fanout(e[:, perm].contiguous(), nodes[:, perm].contiguous()) < N2 or N2 == 1
    TempME._fanout_seen.clear()


def test_executed_flops_accounting():
    """bench.executed_flops: the position-2 event and the [S; P] product once per first-hop slot (walk groups)."""
    import bench
    D, H = 32, 64
    M = H + 12
    full = bench.executed_flops(D, D)
    assert full == 4 * D * D + 12 * D * H + 2 * (2 * H * 3 * H + 2 * H * H + H * M + M * H + H)
    g3 = bench.executed_flops(D, D, fanout=3)
    assert abs((full - g3) - (2.0 / 3.0) * (4 * D * H + 2 * 2 * H * 3 * H)) < 1e-6
    assert bench.executed_flops(D, D, fanout=1) == full
