"""MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:331-400) through hfb_distinctive_descriptors vs the oracle:
representative-descriptor index exact, median within fp32 summation-order tolerance (the difference form is evaluated
literally on both sides; only the order of the 256 additions differs)."""
import numpy as np
import pytest

from hfnet_slam_b200.lib import Context
from oracle import mappoint_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(native_lib):
    with Context(height=64, width=64, n_levels=1, max_keypoints=100, max_batch=1, with_global=False) as c:
        yield c


def _ragged(seed, sizes):
    rng = np.random.default_rng(seed)
    rows, off = [], [0]
    for n in sizes:
        if n:
            centre = rng.standard_normal(256).astype(np.float32)
            d = centre[None] + rng.uniform(0.05, 0.6) * rng.standard_normal((n, 256)).astype(np.float32)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            rows.append(d.astype(np.float32))
        off.append(off[-1] + n)
    desc = np.concatenate(rows) if rows else np.zeros((0, 256), np.float32)
    return desc, np.array(off, np.int32)


def test_matches_oracle_on_ragged_batch(ctx):
    sizes = [1, 2, 3, 5, 8, 13, 30, 0, 64, 17, 2, 128, 7]
    desc, off = _ragged(0, sizes)
    idx, med = ctx.distinctive_descriptors(desc, off)
    ridx, rmed = mappoint_ref.distinctive_batch(desc, off)
    assert np.array_equal(idx, ridx)
    ok = ridx >= 0
    assert np.abs(med[ok] - rmed[ok]).max() < 2e-6
    assert idx[sizes.index(0)] == -1


def test_many_points_and_ties(ctx):
    rng = np.random.default_rng(1)
    sizes = rng.integers(1, 25, 400).tolist()
    desc, off = _ragged(2, sizes)
    # duplicate observations inside a point (the same keyframe feature seen twice): exact zero distances and exact ties
    for p in range(0, 400, 7):
        if sizes[p] >= 3:
            desc[off[p] + 2] = desc[off[p]]
    idx, med = ctx.distinctive_descriptors(desc, off)
    ridx, rmed = mappoint_ref.distinctive_batch(desc, off)
    # medians agree; indices agree wherever the oracle's best median is separated from the runner-up by more than the
    # summation-order noise (exact ties between duplicated rows resolve to the first row on both sides)
    assert np.abs(med - rmed).max() < 2e-6
    mism = np.flatnonzero(idx != ridx)
    for p in mism:
        D = desc[off[p]:off[p + 1]]
        M = np.sort(mappoint_ref.distance_matrix(D), axis=1)[:, int(0.5 * (len(D) - 1))]
        assert abs(M[idx[p]] - M[ridx[p]]) < 2e-6, f"point {p}: {idx[p]} vs {ridx[p]}"
    assert len(mism) <= 2


def test_too_many_observations_is_an_error(ctx):
    from hfnet_slam_b200.lib import HfbError
    desc, off = _ragged(3, [129])
    with pytest.raises(HfbError):
        ctx.distinctive_descriptors(desc, off)
