"""GPU: hardware self-test of the tcgen05 / TMEM conventions (tc.cuh) behind the scorer's GEMMs."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(8, 16), (32, 64), (64, 128), (72, 32), (40, 256), (128, 64)])
def test_selftest_gemm_3xtf32(K, N):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tempme_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(K * 1000 + N)
    A = rng.standard_normal((128, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    dA, dB = torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64)).T
    for mode, tol in ((0, 2e-3), (1, 2e-6)):
        dC = torch.full((128, N), float("nan"), device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.tm_selftest_gemm(_lib.ptr(dA), _lib.ptr(dB), _lib.ptr(dC), K, N, mode, st), "tm_selftest_gemm")
        torch.cuda.synchronize()
        out = dC.cpu().numpy().astype(np.float64)
        err = np.abs(out - ref) / scale
        assert np.isfinite(out).all()
        assert err.max() < tol, (mode, err.max())
    if K % 16 == 0 and N + 2 * K <= 512:      # TS mode: A operand through TMEM (tcgen05.st), same accuracy
        dC = torch.full((128, N), float("nan"), device="cuda")
        _lib.check(L.tm_selftest_gemm(_lib.ptr(dA), _lib.ptr(dB), _lib.ptr(dC), K, N, 3, st), "tm_selftest_gemm")
        torch.cuda.synchronize()
        ts_out = dC.cpu().numpy().astype(np.float64)
        assert (np.abs(ts_out - ref) / scale).max() < 2e-6
    # 3xTF32 must be at fp32-sgemm level: compare with the fp32 product
    f32 = (A @ B.T).astype(np.float64)
    assert np.abs(out - ref).max() <= 4 * np.abs(f32 - ref).max() + 1e-6 * scale.max()


def test_timeencode_cosine_accuracy():
    """The scorer's cosine (exact integer argument reduction) against float64 cos of the same fp32 argument, from
    denormals to the largest finite fp32 (TimeEncode arguments dt * basis_freq reach 1e8, reference explainer.py:55-58)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tempme_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(7)
    mags = 10.0 ** rng.uniform(-12, 38.4, 400000)
    x = np.concatenate([(mags * rng.choice([-1.0, 1.0], mags.size)), rng.uniform(-1e8, 1e8, 400000), rng.uniform(-40, 40, 100000),
                        np.arange(0, 4096) * (np.pi / 2), [0.0, -0.0, 1e-45, 3.4028234e38, -3.4028234e38]]).astype(np.float32)
    dx = torch.as_tensor(x).cuda(); out = torch.empty_like(dx)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.tm_selftest_cos(_lib.ptr(dx), _lib.ptr(out), x.size, st), "tm_selftest_cos")
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    # exact reference: reduce the fp32 argument modulo 2 pi with exact rational arithmetic for the huge ones
    import math
    from fractions import Fraction
    small = np.abs(x) < 1e15
    ref = np.cos(x.astype(np.float64))
    err = np.abs(got - ref)
    assert err[small].max() < 2.5e-7, err[small].max()
    PI = Fraction(3141592653589793238462643383279502884197169399375105820974944592307816406286208998628034825342117067982148086513282306647, 10 ** 120)
    big = np.flatnonzero(~small)[:300]
    for i in big:
        f = Fraction(float(x[i]))
        r = f - (f // (2 * PI)) * (2 * PI)
        assert abs(got[i] - math.cos(float(r))) < 2.5e-7, (x[i], got[i], math.cos(float(r)))
    bad = torch.tensor([float("inf"), float("-inf"), float("nan")], device="cuda"); o2 = torch.zeros(3, device="cuda")
    _lib.check(L.tm_selftest_cos(_lib.ptr(bad), _lib.ptr(o2), 3, st), "tm_selftest_cos")
    assert torch.isnan(o2).all()


@pytest.mark.parametrize("swizzle", [0, 1])
def test_tma_gather4_layout(swizzle):
    """tile::gather4 row gathers (the scorer's feature staging): four table rows per instruction land as consecutive 128-byte
    staging rows; with the 128-byte swizzle the 16-byte chunk c of a row sits at c ^ (bits 7-9 of its shared-memory address);
    columns beyond the table are zero-filled."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tempme_b200 import _lib
    L = _lib.lib()
    rows, dim = 500, 44
    tab = (np.arange(rows)[:, None] * 1000 + np.arange(dim)[None, :]).astype(np.float32)
    idx = np.random.default_rng(0).integers(0, rows, 128).astype(np.int32)
    dt, di = torch.as_tensor(tab).cuda(), torch.as_tensor(idx).cuda()
    for col in (0, 32):
        out = torch.zeros(128 * 32, device="cuda")
        _lib.check(L.tm_selftest_gather4(_lib.ptr(dt), rows, dim, _lib.ptr(di), col, swizzle, _lib.ptr(out), None), "tm_selftest_gather4")
        torch.cuda.synchronize()
        o = out.cpu().numpy().reshape(128, 32)
        shift = int(np.flatnonzero(o[0, ::4] == tab[idx[0], col])[0]) if swizzle else 0     # (address bits 7-9) of staging row 0
        for r in range(128):
            v = np.zeros(32, np.float32); n = max(0, min(32, dim - col)); v[:n] = tab[idx[r], col:col + n]
            x = (r + shift) & 7 if swizzle else 0
            exp = np.concatenate([v[4 * (c ^ x):4 * (c ^ x) + 4] for c in range(8)])
            assert (o[r] == exp).all(), (swizzle, col, r)
