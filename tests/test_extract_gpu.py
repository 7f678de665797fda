"""HF-Net encoder + selection through the C-ABI against the fp32 oracle (oracle/hfnet_ref.py) with seeded synthetic
weights.  Tolerances (fp16 operands / activations, fp32 accumulation; the reference itself runs TensorRT FP16,
src/Extractors/HFNetRTModel.cc:231):  dense score map |err| <= 3e-3 abs, descriptor map cosine >= 0.999,
global descriptor cosine >= 0.999; keypoint selection is bit-exact GIVEN the device's own dense maps (re-run through the
oracle's post-processing) and >= 90 % identical to the all-fp32 oracle selection."""
import numpy as np
import pytest

from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context
from oracle import hfnet_ref, select_ref

pytestmark = pytest.mark.gpu
SCORE_TOL = 3e-3
COS_TOL = 0.999


@pytest.fixture(scope="module")
def ctx_euroc(native_lib, weights_blob):
    import os
    os.environ["HFB_DEBUG"] = "1"
    ctx = Context(height=480, width=752, n_levels=1, max_keypoints=1000, max_batch=2, with_global=True)
    ctx.load_weights(weights_blob)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def oracle_euroc(weights_dict):
    img = weights.synthetic_image(480, 752, seed=1)
    return img, hfnet_ref.forward(img, weights_dict, want_global=True, return_intermediates=True)


def _cos(a, b):
    a, b = a.reshape(-1, a.shape[-1]).astype(np.float64), b.reshape(-1, b.shape[-1]).astype(np.float64)
    return (a * b).sum(1) / np.maximum(np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1), 1e-30)


def test_layers_track_oracle(ctx_euroc, oracle_euroc):
    img, ref = oracle_euroc
    ctx_euroc.extract(img, [1000], 0.01)
    report = []
    for name in ["layer_1", "layer_2", "layer_3", "layer_4", "layer_5", "layer_6", "layer_7", "layer_8", "layer_12",
                 "layer_15", "layer_18", "desc_conv1", "det_conv1", "det_logits"]:
        got = ctx_euroc.debug_tensor(name)[0]
        r = ref[name][0]
        assert got.shape == r.shape, f"{name}: shape {got.shape} vs {r.shape}"
        scale = np.abs(r).max() + 1e-9
        err = np.abs(got - r).max() / scale
        report.append((name, float(err)))
    msg = ", ".join(f"{n}:{e:.2e}" for n, e in report)
    # fp16 activations: rounding error compounds with depth (observed 3e-4 at layer_1 ... 5e-2 at layer_18, max-norm)
    lim = lambda n: 0.1 if n in ("layer_12", "layer_15", "layer_18") else 3e-2
    assert all(e < lim(n) for n, e in report), "relative max error per layer: " + msg


def test_dense_outputs(ctx_euroc, oracle_euroc):
    img, ref = oracle_euroc
    ctx_euroc.extract(img, [1000], 0.01)
    scores = ctx_euroc.debug_tensor("scores_dense")[0, :, :, 0]
    err = np.abs(scores - ref["scores_dense"][0]).max()
    assert err <= SCORE_TOL, f"dense score map max abs err {err}"
    dm = ctx_euroc.debug_tensor("local_descriptor_map")[0]
    c = _cos(dm, ref["local_descriptor_map"][0])
    assert c.min() >= COS_TOL, f"descriptor map min cosine {c.min()}"
    assert np.abs(np.linalg.norm(dm, axis=-1) - 1).max() < 1e-5
    g = ctx_euroc.debug_tensor("global_descriptor").reshape(1, -1)
    cg = _cos(g, ref["global_descriptor"])
    assert cg.min() >= COS_TOL, f"global descriptor cosine {cg.min()}"


def test_selection_exact_on_device_maps(ctx_euroc, oracle_euroc):
    img, ref = oracle_euroc
    out = ctx_euroc.extract(img, [1000], 0.01)
    scores = ctx_euroc.debug_tensor("scores_dense")[0, :, :, 0]
    nms_dev = ctx_euroc.debug_tensor("scores_dense_nms")[0, :, :, 0]
    import torch
    nms_ref = hfnet_ref.simple_nms(torch.from_numpy(scores)[None], 4, 2)[0].numpy()
    assert np.array_equal(nms_dev, nms_ref), "in-graph NMS differs from the oracle on the device's own score map"
    dm = ctx_euroc.debug_tensor("local_descriptor_map")[0]
    exp = select_ref.local_features(nms_dev, dm, 1000, 0.01)
    assert out["n_per_level"][0] == len(exp["x"]) == len(out["x"])
    for k in ("x", "y", "response"):
        assert np.array_equal(out[k], exp[k]), k
    assert np.array_equal(out["descriptors"], exp["descriptors"])
    assert (out["octave"] == 0).all()
    # against the all-fp32 oracle: same keypoints up to score-noise swaps at the cut
    full = select_ref.local_features(ref["scores_dense_nms"][0], ref["local_descriptor_map"][0], 1000, 0.01)
    a = set(zip(out["x"].astype(int).tolist(), out["y"].astype(int).tolist()))
    b = set(zip(full["x"].astype(int).tolist(), full["y"].astype(int).tolist()))
    iou = len(a & b) / max(len(a | b), 1)
    assert len(b) > 100, "synthetic weights should give a non-trivial keypoint set"
    assert iou >= 0.9, f"keypoint set IoU vs fp32 oracle {iou:.3f} ({len(a)} vs {len(b)})"
    g = out["global_descriptor"].reshape(1, -1)
    assert _cos(g, ref["global_descriptor"]).min() >= COS_TOL


def test_batch_equals_single(ctx_euroc, oracle_euroc):
    img, _ = oracle_euroc
    img2 = weights.synthetic_image(480, 752, seed=5)
    one = ctx_euroc.extract(img, [1000], 0.01)
    two = ctx_euroc.extract_batch([img2, img], [1000], 0.01)
    for k in ("x", "y", "response", "descriptors", "global_descriptor"):
        assert np.array_equal(one[k], two[1][k]), k
    assert not np.array_equal(two[0]["global_descriptor"], two[1]["global_descriptor"])


@pytest.mark.parametrize("H,W,n_feat,thr", [(240, 376, 675, 0.01), (512, 512, 850, 0.02)],
                         ids=["euroc-4level-small", "tumvi-512x512-4level"])
def test_multilevel_pyramid(native_lib, weights_blob, weights_dict, H, W, n_feat, thr):
    """4-level configurations: EuRoC (Examples/Monocular/EuRoC.yaml:67-80 -> 675 features, 1.2, 4 levels; smaller frame)
    and the TUM-VI shape of BASELINE.json configs[4] (Examples/Monocular/TUM-VI.yaml:66-67 -> 850 features -> 274 / 228 /
    190 / 158 per level, threshold 0.02, 512 x 512)."""
    img = weights.synthetic_image(H, W, seed=2, n_corners=80)
    budgets = select_ref.features_per_level(n_feat, 4, 1.2)
    if n_feat == 850:
        assert budgets == [274, 228, 190, 158]
    with Context(height=H, width=W, n_levels=4, scale_factor=1.2, max_keypoints=1000, max_batch=1) as ctx:
        ctx.load_weights(weights_blob)
        out = ctx.extract(img, budgets, thr)
        pyr = select_ref.compute_pyramid(img, 4, 1.2)
        per_level = []
        for l, im in enumerate(pyr):
            nms = ctx.debug_tensor("scores_dense_nms", level=l)[0, :, :, 0]
            dm = ctx.debug_tensor("local_descriptor_map", level=l)[0]
            assert nms.shape == (im.shape[0] // 8 * 8, im.shape[1] // 8 * 8)
            per_level.append(select_ref.local_features(nms, dm, budgets[l], thr))
            # dense maps of every level track the fp32 oracle run on the cv2 pyramid
            r = hfnet_ref.forward(im, weights_dict, want_global=False)
            sc = ctx.debug_tensor("scores_dense", level=l)[0, :, :, 0]
            assert np.abs(sc - r["scores_dense"][0]).max() <= SCORE_TOL, f"level {l}"
        exp = select_ref.concat_levels(per_level, 1.2)
        assert out["n_per_level"][:4] == [len(p["x"]) for p in per_level]
        for k in ("x", "y", "response", "octave", "descriptors"):
            assert np.array_equal(out[k], exp[k]), k


def test_errors_are_loud(native_lib, weights_blob):
    from hfnet_slam_b200.lib import HfbError
    with Context(height=64, width=64, n_levels=1, max_keypoints=100, max_batch=1) as ctx:
        with pytest.raises(HfbError):
            ctx.extract(np.zeros((64, 64), np.uint8), [10], 0.01)          # weights not loaded
        ctx.load_weights(weights_blob)
        with pytest.raises(HfbError):
            ctx.extract(np.zeros((32, 64), np.uint8), [10], 0.01)          # wrong shape
        with pytest.raises(HfbError):
            ctx.extract(np.zeros((64, 64), np.uint8), [1000], 0.01)        # budget above capacity
        with pytest.raises(HfbError):
            ctx.load_weights(b"garbage" * 10)
        out = ctx.extract(np.zeros((64, 64), np.uint8), [0], 0.01)         # zero budget is legal
        assert out["x"].size == 0


def test_pinned_zero_copy_path_equals_staged_path(ctx_euroc, oracle_euroc):
    """Frames / outputs in page-locked memory (hfb_host_alloc) are DMA'd in place; results must be identical."""
    from hfnet_slam_b200.lib import pinned_empty
    img, _ = oracle_euroc
    img2 = weights.synthetic_image(480, 752, seed=9)
    staged = ctx_euroc.extract_batch([img, img2], [1000], 0.01)
    pin = []
    for im in (img, img2):
        p = pinned_empty(im.shape, np.uint8)
        p[...] = im
        pin.append(p)
    direct, block = ctx_euroc.extract_batch(pin, [1000], 0.01, return_block=True, pinned=True)
    for a, b in zip(staged, direct):
        for k in ("x", "y", "response", "octave", "descriptors", "global_descriptor"):
            assert np.array_equal(a[k], b[k]), k
    assert block["descriptors"].shape == (2, 1000, 256)


def test_consecutive_match_on_resident_descriptors(ctx_euroc):
    """hfb_match_consecutive: frame b vs frame b-1 on the descriptors the extraction left in HBM == the oracle's
    cv::BFMatcher(NORM_L2, crossCheck) + dist < 0.6 (src/Matcher.cc:220-263) on the descriptors returned to the host."""
    from oracle import match_ref
    base = weights.synthetic_image(480, 752, seed=7)
    imgs = [base, np.roll(base, (3, 5), axis=(0, 1))]
    feats = ctx_euroc.extract_batch(imgs, [1000], 0.01)
    idx, val = ctx_euroc.match_consecutive(2, 0, 0.6)
    assert idx.shape == (2, ctx_euroc.kp_cap)
    for b in range(2):
        A, Bd = feats[b]["descriptors"], feats[(b - 1) % 2]["descriptors"]
        ia, ib, dist = match_ref.search_by_bow(A, Bd, 0.6)
        got = {(int(i), int(idx[b, i])) for i in np.flatnonzero(idx[b, :len(A)] >= 0)}
        want = set(zip(ia.tolist(), ib.tolist()))
        # identical sets, except pairs sitting on the decision boundaries within the fp32 noise of the two distance
        # evaluations (|dist - TH_LOW| or the gap to the runner-up below 5e-6)
        for i, j in got ^ want:
            dm = match_ref.l2_distance_matrix(A[i:i + 1], Bd)[0]
            dcol = match_ref.l2_distance_matrix(A, Bd[j:j + 1])[:, 0]
            gap_row = np.partition(dm, 1)[1] - np.partition(dm, 1)[0]
            gap_col = np.partition(dcol, 1)[1] - np.partition(dcol, 1)[0]
            assert abs(dm[j] - 0.6) < 5e-6 or gap_row < 5e-6 or gap_col < 5e-6, f"frame {b}: pair {(i, j)} differs"
        assert len(got ^ want) <= 2, f"frame {b}"
        assert len(got) > 50, "shifted copies of one frame should share many keypoints"
        one_idx, one_val = ctx_euroc.fetch_matches(b, len(A))
        assert np.array_equal(one_idx, idx[b, :len(A)]) and np.array_equal(one_val, val[b, :len(A)])
        assert (idx[b, len(A):] < 0).all()
    # the one-call form (association enqueued inside the extraction, ahead of the global branch) gives the same rows,
    # through pageable and through page-locked outputs
    for pinned in (False, True):
        feats2, idx2, val2 = ctx_euroc.extract_match_batch(imgs, [1000], 0.01, 0, 0.6, pinned=pinned)
        assert np.array_equal(idx2, idx) and np.array_equal(val2, val), f"pinned={pinned}"
        for b in range(2):
            for k in ("x", "y", "response", "descriptors", "global_descriptor"):
                assert np.array_equal(feats2[b][k], feats[b][k]), (pinned, b, k)
