"""The bf16x3 tensor-core GEMM (rb_gemm_bf16x3) vs float64 numpy: every operand orientation, ragged shapes, epilogues, split-K."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, a_mn, B, b_mn, M, N, K, bias=None, act=0, C0=None, H=None, ws_floats=0):
    from reacherdistilation_b200._lib import check, lib, ptr, stream_ptr
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    C = torch.from_numpy(C0.copy()).cuda() if C0 is not None else torch.full((M, N), float("nan"), device="cuda")
    db = torch.from_numpy(bias).cuda() if bias is not None else None
    dH = torch.from_numpy(H).cuda() if H is not None else None
    ws = torch.empty(ws_floats, device="cuda") if ws_floats else None
    check(lib().rb_gemm_bf16x3(ptr(dA), A.shape[1], a_mn, ptr(dB), B.shape[1], b_mn, ptr(C), N, M, N, K, ptr(db), act, 1 if C0 is not None else 0,
                               ptr(dH), N if H is not None else 0, ptr(ws), ws_floats, stream_ptr()))
    return C.cpu().numpy()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1, 4, 4), (20, 800, 243), (300, 64, 200), (129, 33, 17), (4096, 32, 64), (243, 800, 1000), (4, 32, 5000)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
def test_gemm_orientations(M, N, K, a_mn, b_mn):
    rng = np.random.default_rng(M + 7 * N + 13 * K + a_mn + 2 * b_mn)
    Am, Bm = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)      # logical A [M,K], B [K,N]
    A = np.ascontiguousarray(Am.T) if a_mn else Am                 # a_mn = 1: stored [K][M]
    B = Bm if b_mn else np.ascontiguousarray(Bm.T)                 # b_mn = 1: stored [K][N]; 0: stored [N][K]
    ref = Am.astype(np.float64) @ Bm.astype(np.float64)
    ws = 32 * M * N if K >= 1000 else 0
    out = _run(A, a_mn, B, b_mn, M, N, K, ws_floats=ws)
    err = np.abs(out - ref).max() / max(1.0, np.abs(ref).max())
    assert err <= 2e-5, err


def test_gemm_epilogues_and_accumulate():
    rng = np.random.default_rng(0)
    M, N, K = 200, 96, 72
    Am, Bm = rng.standard_normal((M, K)).astype(np.float32) * 0.3, rng.standard_normal((K, N)).astype(np.float32) * 0.3
    bias, H, C0 = rng.standard_normal(N).astype(np.float32), np.tanh(rng.standard_normal((M, N))).astype(np.float32), rng.standard_normal((M, N)).astype(np.float32)
    ref = Am.astype(np.float64) @ Bm.astype(np.float64)
    rel = lambda out, r: (np.abs(out - r) / np.maximum(1.0, np.abs(r))).max()
    assert rel(_run(Am, 0, Bm, 1, M, N, K, bias=bias, act=1), np.tanh(ref + bias)) <= 5e-5
    assert rel(_run(Am, 0, Bm, 1, M, N, K, H=H), ref * (1 - H.astype(np.float64) ** 2)) <= 5e-5
    assert rel(_run(Am, 0, Bm, 1, M, N, K, C0=C0), ref + C0) <= 5e-5
    a = _run(Am, 0, Bm, 1, M, N, K, bias=bias)
    assert np.array_equal(a, _run(Am, 0, Bm, 1, M, N, K, bias=bias))                     # deterministic


@pytest.mark.parametrize("M,N,K,ws", [(300, 128, 64, 0), (257, 200, 200, 0), (243, 800, 1000, 32 * 243 * 800), (1024, 64, 200, 0), (640, 32, 64, 0), (512, 4, 32, 0), (128, 65, 7, 0)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (0, 0)])
def test_cta_packing_is_bit_identical(M, N, K, ws, a_mn, b_mn):
    """Deep pipeline (2 bf16 + 2-4 raw stages) vs packed CTAs (1 + 2, two / three CTAs per SM): the same MMAs in the same order, every tile width."""
    from reacherdistilation_b200._lib import check, lib
    rng = np.random.default_rng(M + N + K)
    sc = 0.1 / np.sqrt(max(K, 64) / 64.0)                              # pre-activations of order one (the bf16x3 error scales with them)
    Am, Bm = (rng.standard_normal((M, K)) * sc).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    A = np.ascontiguousarray(Am.T) if a_mn else Am
    B = Bm if b_mn else np.ascontiguousarray(Bm.T)
    bias = rng.standard_normal(N).astype(np.float32)
    outs = []
    try:
        for mode in (0, 1):
            check(lib().rb_gemm_set_cta_packing(mode))
            outs.append(_run(A, a_mn, B, b_mn, M, N, K, bias=bias, act=1, ws_floats=ws))
    finally:
        check(lib().rb_gemm_set_cta_packing(-1))
    ref = np.tanh(Am.astype(np.float64) @ Bm.astype(np.float64) + bias)
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[1] - ref).max() <= 5e-5
