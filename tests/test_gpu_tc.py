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
