"""Matcher kernels through the C-ABI against oracle/match_ref.py: identical match sets (bit-exact indices), values within
2e-6 (fp32 summation-order noise, written here)."""
import numpy as np
import pytest

from hfnet_slam_b200 import synthetic
from oracle import match_ref

pytestmark = pytest.mark.gpu
VAL_TOL = 2e-6


def _check(idx, val, ref, na):
    ia, ib, v = ref
    got = {(int(i), int(idx[i])) for i in range(na) if idx[i] >= 0}
    exp = set(zip(ia.tolist(), ib.tolist()))
    assert got == exp, f"match sets differ: missing {sorted(exp - got)[:5]} extra {sorted(got - exp)[:5]} ({len(got)} vs {len(exp)})"
    if len(ia):
        assert np.abs(val[ia] - v).max() <= VAL_TOL


@pytest.mark.parametrize("na,nb,seed", [(1000, 1000, 0), (675, 850, 1), (1, 1, 2), (5, 300, 3), (129, 127, 4), (2000, 1500, 5)])
def test_mutual_l2(small_ctx, na, nb, seed):
    A, B = synthetic.descriptor_pair(na, nb, seed=seed, n_true=min(300, nb, max(na - 100, 0)))
    idx, val, n = small_ctx.match_mutual_l2(A, B, 0.6)
    _check(idx, val, match_ref.search_by_bow(A, B, 0.6), na)
    assert n == int((idx >= 0).sum())


@pytest.mark.parametrize("na,nb,seed", [(1000, 1000, 0), (700, 333, 7), (130, 1000, 8)])
def test_mutual_cos(small_ctx, na, nb, seed):
    A, B = synthetic.descriptor_pair(na, nb, seed=seed, n_true=min(300, nb, max(na - 100, 0)))
    idx, val, n = small_ctx.match_mutual_cos(A, B, float(match_ref.COS_FLOOR))
    _check(idx, val, match_ref.search_for_triangulation_core(A, B), na)


def test_empty_inputs(small_ctx):
    A, B = synthetic.descriptor_pair(10, 10, n_true=0)
    idx, val, n = small_ctx.match_mutual_l2(A[:0], B, 0.6)
    assert idx.size == 0 and n == 0
    idx, val, n = small_ctx.match_mutual_l2(A, B[:0], 0.6)
    assert (idx == -1).all() and n == 0


def test_duplicate_rows_lowest_index_wins(small_ctx):
    A, B = synthetic.descriptor_pair(64, 64, n_true=0, seed=11)
    B[10] = A[3]
    B[20] = A[3]          # exact tie on the row: the lower column index must win (strict '>' scan, Matcher.cc:868)
    A[40] = A[3]          # exact tie on the column: the lower row index must win
    idx, val, _ = small_ctx.match_mutual_cos(A, B, float(match_ref.COS_FLOOR))
    _check(idx, val, match_ref.search_for_triangulation_core(A, B), 64)
    assert idx[3] == 10 and idx[40] == -1


def test_batched_pairs_ragged(small_ctx):
    rng = np.random.default_rng(5)
    cnt_a, cnt_b = [300, 0, 129, 700, 1], [250, 40, 0, 650, 5]
    As, Bs = [], []
    for i, (a, b) in enumerate(zip(cnt_a, cnt_b)):
        A, B = synthetic.descriptor_pair(max(a, 1), max(b, 1), seed=20 + i, n_true=min(100, b, max(a - 100, 0)))
        As.append(A[:a]); Bs.append(B[:b])
    A_all, B_all = np.concatenate(As), np.concatenate(Bs)
    a_off = np.concatenate([[0], np.cumsum(cnt_a)[:-1]]).astype(np.int32)
    b_off = np.concatenate([[0], np.cumsum(cnt_b)[:-1]]).astype(np.int32)
    for mode, thr, ref_fn in ((0, 0.6, lambda a, b: match_ref.search_by_bow(a, b, 0.6)),
                              (1, float(match_ref.COS_FLOOR), match_ref.search_for_triangulation_core)):
        idx, val = small_ctx.match_batch(mode, A_all, B_all, a_off, cnt_a, b_off, cnt_b, thr)
        for p in range(len(cnt_a)):
            sl = slice(a_off[p], a_off[p] + cnt_a[p])
            _check(idx[sl], val[sl], ref_fn(As[p], Bs[p]), cnt_a[p])


def test_non_unit_rows_l2(small_ctx):
    """BFMatcher does not assume unit norm: scale rows and check the L2 mutual-NN set still matches."""
    A, B = synthetic.descriptor_pair(400, 380, seed=9, n_true=200)
    rng = np.random.default_rng(1)
    A = (A * rng.uniform(0.8, 1.2, (400, 1))).astype(np.float32)
    B = (B * rng.uniform(0.8, 1.2, (380, 1))).astype(np.float32)
    idx, val, _ = small_ctx.match_mutual_l2(A, B, 0.6)
    _check(idx, val, match_ref.search_by_bow(A, B, 0.6), 400)


def test_keyframe_store_neighbour_matching_equals_batch_call(native_lib):
    """hfb_match_kf_neighbours on descriptors RESIDENT in the keyframe store == hfb_match_batch on host copies of the same
    descriptors (both flavours), also across two contexts (store filled by one, read by the other) and after erase / reuse
    of a slot."""
    from hfnet_slam_b200.lib import Context, KeyFrameStore
    rng = np.random.default_rng(4)
    base = rng.standard_normal((700, 256)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    kfs = []
    for k, n in enumerate((675, 640, 700, 1, 333, 675)):
        d = base[rng.permutation(700)[:n]] + 0.03 * rng.standard_normal((n, 256)).astype(np.float32)
        kfs.append((d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32))
    with Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False) as ca, \
            Context(height=64, width=64, n_levels=1, max_keypoints=64, max_batch=1, with_global=False) as cb:
        store = KeyFrameStore(ca, n_slots=6, rows_per_slot=700)
        for k, d in enumerate(kfs):
            store.put(ca, 10 + k, d)
        assert len(store) == 6
        with pytest.raises(Exception):
            store.put(ca, 99, kfs[0])                      # full
        A = kfs[0]
        nb = [1, 2, 3, 4, 5]
        for mode, thr in ((0, 0.6), (1, 0.71875)):
            idx, val = store.match_neighbours(cb, 10, [10 + k for k in nb], mode, thr, len(A))
            A_all = np.concatenate([A] * len(nb))
            B_all = np.concatenate([kfs[k] for k in nb])
            cnt_b = np.array([len(kfs[k]) for k in nb], np.int32)
            ref_i, ref_v = cb.match_batch(mode, A_all, B_all, (np.arange(len(nb)) * len(A)).astype(np.int32),
                                          np.full(len(nb), len(A), np.int32), (np.cumsum(cnt_b) - cnt_b).astype(np.int32), cnt_b, thr)
            assert np.array_equal(idx.reshape(-1), ref_i) and np.array_equal(val.reshape(-1), ref_v), mode
            assert (idx[0] >= 0).sum() > 100
        store.erase(12)
        store.put(cb, 50, kfs[2][:100])                    # the freed slot is reused, shorter than before
        idx, val = store.match_neighbours(ca, 10, [50], 0, 0.6, len(A))
        ref_i, ref_v, _ = ca.match_mutual_l2(A, kfs[2][:100], 0.6)
        assert np.array_equal(idx[0], ref_i) and np.array_equal(val[0], ref_v)
        store.close()
