"""tcgen05/TMA GEMM primitive (hfnet_slam_b200/csrc/gemm_core.cuh) against numpy on fp16-rounded operands and
against the CUDA-core restatement on the device.  Tolerance: fp32 accumulation of fp16 products, |err| <= 2e-3 * sqrt(K)
relative to unit-scale operands (written below)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(A, Wt, bias, relu6):
    a = A.astype(np.float16).astype(np.float64)
    w = Wt.astype(np.float16).astype(np.float64)
    out = a @ w.T + (bias.astype(np.float64) if bias is not None else 0.0)
    if relu6:
        out = np.clip(out, 0.0, 6.0)
    return out.astype(np.float32)


PLAIN = [(128, 16, 64), (300, 24, 16), (1000, 96, 24), (257, 144, 24), (129, 288, 96), (360, 720, 240),
         (128, 256, 256), (5640, 128, 80), (77, 64, 32), (4096, 768, 128)]


@pytest.mark.parametrize("M,K,N", PLAIN)
def test_plain_gemm_matches_numpy(small_ctx, M, K, N):
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    A = rng.normal(size=(M, K)).astype(np.float32)
    Wt = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    got = small_ctx.debug_gemm(A, Wt, bias=bias, relu6=False, use_tc=True)
    ref = _ref(A, Wt, bias, False)
    simt = small_ctx.debug_gemm(A, Wt, bias=bias, relu6=False, use_tc=False)
    err, err_simt = np.abs(got - ref).max(), np.abs(simt - ref).max()
    assert err_simt < 1e-3, f"CUDA-core restatement off by {err_simt}"
    bad = np.argwhere(np.abs(got - ref) > 1e-3)
    assert err < 1e-3, (f"tcgen05 GEMM M={M} K={K} N={N}: max err {err}, {len(bad)} bad entries, first {bad[:8].tolist()}, "
                        f"got {got[tuple(bad[0])] if len(bad) else None} ref {ref[tuple(bad[0])] if len(bad) else None}")


@pytest.mark.parametrize("bn", [16, 48, 96, 144, 256])
def test_plain_gemm_n_tiles(small_ctx, bn):
    rng = np.random.default_rng(bn)
    M, K, N = 700, 144, 288
    A = rng.normal(size=(M, K)).astype(np.float32)
    Wt = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    got = small_ctx.debug_gemm(A, Wt, relu6=True, use_tc=True, BN=bn)
    ref = _ref(A, Wt, None, True)
    assert np.abs(got - ref).max() < 1e-3


def test_gemm_identity_pattern(small_ctx):
    """A = shifted identity rows: C[m][n] = W[n][m % K] -- any swizzle / descriptor mistake permutes K visibly."""
    M, K, N = 256, 128, 64
    A = np.zeros((M, K), np.float32)
    A[np.arange(M), np.arange(M) % K] = 1.0
    Wt = (np.arange(N * K, dtype=np.float32).reshape(N, K) % 251) / 64.0
    got = small_ctx.debug_gemm(A, Wt, use_tc=True)
    ref = Wt.astype(np.float16).astype(np.float32).T[np.arange(M) % K]
    bad = np.argwhere(got != ref)
    assert len(bad) == 0, f"{len(bad)} mismatches, first {bad[:10].tolist()}"


@pytest.mark.parametrize("B,H,W,C,N", [(1, 8, 16, 64, 32), (1, 60, 94, 96, 384), (2, 13, 21, 96, 128), (1, 9, 17, 32, 16)])
def test_conv3x3_implicit_gemm(small_ctx, B, H, W, C, N):
    rng = np.random.default_rng(H * W + C)
    X = rng.normal(size=(B, H, W, C)).astype(np.float32)
    Wt = (rng.normal(size=(N, 9 * C)) / np.sqrt(9 * C)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    got = small_ctx.debug_gemm(X, Wt, bias=bias, relu6=True, conv3x3=True, use_tc=True, BN=min(128, (N + 15) // 16 * 16))
    xp = np.pad(X.astype(np.float16).astype(np.float64), ((0, 0), (1, 1), (1, 1), (0, 0)))
    w = Wt.astype(np.float16).astype(np.float64).reshape(N, 9, C)
    ref = np.zeros((B, H, W, N))
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        ref += xp[:, dy:dy + H, dx:dx + W, :] @ w[:, tap, :].T
    ref = np.clip(ref + bias, 0, 6).astype(np.float32)
    simt = small_ctx.debug_gemm(X, Wt, bias=bias, relu6=True, conv3x3=True, use_tc=False)
    assert np.abs(simt - ref).max() < 1e-3
    err = np.abs(got - ref)
    bad = np.argwhere(err > 1e-3)
    assert err.max() < 1e-3, f"conv3x3 max err {err.max()} ({len(bad)} bad, first {bad[:6].tolist()})"
