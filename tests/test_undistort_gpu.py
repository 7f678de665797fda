"""Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:760-825) through the C-ABI: bit-exact against the oracle
(itself pinned bit for bit to cv2.undistortPoints in tests/test_oracle_pins.py), on caller-supplied points, on the resident
keypoints of an extraction, and as the coordinates the resident windowed search uses."""
import numpy as np
import pytest

from hfnet_slam_b200 import weights
from hfnet_slam_b200.lib import Context
from oracle import select_ref

pytestmark = pytest.mark.gpu

# EuRoC cam0 (Examples/Monocular/EuRoC.yaml:9-22) and a 5- / 8- / 12-coefficient vector
CAMS = [((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
        ((190.978, 190.973, 254.932, 256.897), (-0.1, 0.02, 0.001, -0.0005, 0.003)),
        ((300.0, 300.0, 320.0, 240.0), (-0.9, 0.1, 0.01, 0.01, 0.0, 0.3, 0.01, 0.002)),          # reaches icdist < 0
        ((300.0, 300.0, 320.0, 240.0), (-0.4, 0.1, 0.01, 0.01, 0.0, 0.3, 0.01, 0.002, 1e-3, 2e-3, -1e-3, 5e-4))]


def _same(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("cam", range(len(CAMS)))
@pytest.mark.parametrize("n", [1, 4, 1000, 4097])
def test_undistort_points_bit_exact(small_ctx, cam, n):
    K, dist = CAMS[cam]
    rng = np.random.default_rng(10 * cam + n)
    x = rng.uniform(-40, 800, n).astype(np.float32)
    y = rng.uniform(-40, 520, n).astype(np.float32)
    small_ctx.set_camera(K, dist)
    with np.errstate(all="ignore"):
        ex, ey = select_ref.undistort_points(x, y, K, dist)
    gx, gy = small_ctx.undistort_points(x, y)
    assert _same(gx, ex) and _same(gy, ey)
    with np.errstate(all="ignore"):
        assert np.array_equal(small_ctx.image_bounds(752, 480), select_ref.image_bounds(752, 480, K, dist), equal_nan=True)


def test_zero_distortion_is_identity(small_ctx):
    """dist[0] == 0 is the reference's early return, whatever the other coefficients say (src/Frame.cc:762-766)."""
    x = np.arange(50, dtype=np.float32) * 3.7
    y = np.arange(50, dtype=np.float32) * 1.3
    for dist in ((), (0.0, 0.5, 0.1, 0.1)):
        small_ctx.set_camera((400, 400, 320, 240), dist)
        gx, gy = small_ctx.undistort_points(x, y)
        assert np.array_equal(gx, x) and np.array_equal(gy, y)
        assert small_ctx.image_bounds(640, 480).tolist() == [0.0, 640.0, 0.0, 480.0]
    gx, gy = small_ctx.undistort_points(x[:0], y[:0])
    assert gx.size == 0
    with pytest.raises(RuntimeError):
        small_ctx.set_camera((400, 400, 320, 240), (0.1, 0.2, 0.3))          # not an OpenCV coefficient count


def test_resident_keypoints_are_undistorted_in_the_extraction(native_lib, weights_blob):
    """With a distorted camera every extraction (here 2 levels, batch of 2, graph replay included) leaves mvKeysUn next
    to mvKeys; changing the calibration re-captures; the resident windowed search sees the undistorted coordinates."""
    K, dist = CAMS[0]
    H, W = 240, 376
    imgs = [weights.synthetic_image(H, W, seed=s, n_corners=60) for s in (1, 2)]
    budgets = select_ref.features_per_level(300, 2, 1.2)
    with Context(height=H, width=W, n_levels=2, scale_factor=1.2, max_keypoints=512, max_batch=2) as ctx:
        ctx.load_weights(weights_blob)
        plain = ctx.extract_batch(imgs, budgets, 0.01)
        px, py = ctx.fetch_undistorted(1, len(plain[1]["x"]))
        assert np.array_equal(px, plain[1]["x"]) and np.array_equal(py, plain[1]["y"])          # no camera: mvKeysUn = mvKeys
        ctx.set_camera(K, dist)
        with pytest.raises(RuntimeError):
            ctx.fetch_undistorted(0, 1)                                                         # extracted before the camera was set
        for _ in range(3):                                                                      # warm run, capture, replay
            feats = ctx.extract_batch(imgs, budgets, 0.01)
            for b in range(2):
                n = len(feats[b]["x"])
                assert n > 50 and np.array_equal(feats[b]["x"], plain[b]["x"])
                ex, ey = select_ref.undistort_points(feats[b]["x"], feats[b]["y"], K, dist)
                gx, gy = ctx.fetch_undistorted(b, n)
                assert _same(gx, ex) and _same(gy, ey)
        # resident search: a window centred on an undistorted keypoint of frame 1 finds that keypoint; centred on its
        # distorted position (several pixels away near the border) with the same radius it does not
        n1 = len(feats[1]["x"])
        ex, ey = select_ref.undistort_points(feats[1]["x"], feats[1]["y"], K, dist)
        shift = np.hypot(ex - feats[1]["x"], ey - feats[1]["y"])
        j = int(np.argmax(shift))
        assert shift[j] > 2.0
        q = np.array([0], np.int32)
        for centre, hit in (((ex[j], ey[j]), True), ((feats[1]["x"][j], feats[1]["y"][j]), False)):
            idx, _, _ = ctx.match_projection_frame(1, q, np.array([centre], np.float32), np.array([0.75], np.float32),
                                                   np.array([0], np.int32), np.array([-1], np.int32), n1)
            assert (j in idx[0].tolist()) == hit
