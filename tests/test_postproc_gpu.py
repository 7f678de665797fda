"""Network-tail kernels through the C-ABI, bit-exact against the oracle on identical dense maps:
simple_nms, threshold scan + top-k + Resampler + normalize, cv::resize(INTER_LINEAR) u8."""
import numpy as np
import pytest
import torch

from oracle import hfnet_ref, select_ref

pytestmark = pytest.mark.gpu


def _score_map(h, w, seed, plateau=False):
    rng = np.random.default_rng(seed)
    s = rng.random((h, w), dtype=np.float32) ** 6
    if plateau:                       # equal neighbours: both survive NMS, exercises ties
        s[10:14, 20:23] = 0.9
        s[h - 1, w - 3:] = 0.95
        s[0, 0] = 1.0
    return s


@pytest.mark.parametrize("h,w,seed,plateau", [(480, 752, 0, False), (64, 64, 1, True), (272, 432, 2, True), (17, 9, 3, False),
                                                (8, 200, 4, True)])
def test_nms_bit_exact(small_ctx, h, w, seed, plateau):
    s = _score_map(h, w, seed, plateau)
    ref = hfnet_ref.simple_nms(torch.from_numpy(s)[None], 4, 2)[0].numpy()
    got = small_ctx.nms(s)
    bad = np.argwhere(got != ref)
    assert len(bad) == 0, f"{len(bad)} pixels differ, first {bad[:8].tolist()}"


@pytest.mark.parametrize("h,w,k,thr,seed", [(480, 752, 1000, 0.01, 0), (400, 624, 181, 0.01, 1), (64, 64, 8192, 0.0005, 2),
                                              (64, 96, 50, 2.0, 3), (128, 128, 300, 0.3, 4)])
def test_select_sample_bit_exact(small_ctx, h, w, k, thr, seed):
    rng = np.random.default_rng(seed)
    s = _score_map(h, w, seed, plateau=True)
    nms = hfnet_ref.simple_nms(torch.from_numpy(s)[None], 4, 2)[0].numpy()
    nms[5, 7] = nms[9, 3] = nms[9, 4] = 0.77          # exact response ties across the cut order
    dm = rng.normal(size=(h // 8, w // 8, 256)).astype(np.float32)
    dm /= np.linalg.norm(dm, axis=-1, keepdims=True)
    ref = select_ref.local_features(nms, dm, k, thr)
    got = small_ctx.select_sample(nms, dm, k, thr)
    assert len(got["x"]) == len(ref["x"]), f"{len(got['x'])} vs {len(ref['x'])} keypoints"
    for key in ("x", "y", "response"):
        assert np.array_equal(got[key], ref[key]), f"{key} differs (first bad {np.flatnonzero(got[key] != ref[key])[:5]})"
    d = np.abs(got["descriptors"] - ref["descriptors"])
    assert np.array_equal(got["descriptors"], ref["descriptors"]), f"descriptors differ: max {d.max()}, rows {np.unique(np.argwhere(d > 0)[:, 0])[:8]}"


@pytest.mark.parametrize("sh,sw,dh,dw", [(480, 752, 400, 627), (400, 627, 333, 522), (333, 522, 278, 435), (512, 512, 427, 427),
                                           (37, 53, 31, 44)])
def test_resize_bit_exact(small_ctx, sh, sw, dh, dw):
    import cv2
    rng = np.random.default_rng(sh + dw)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    got = small_ctx.resize_linear_u8(src, dh, dw)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    bad = np.argwhere(got != ref)
    assert len(bad) == 0, f"{len(bad)} pixels differ from cv2.resize, first {bad[:8].tolist()}"
