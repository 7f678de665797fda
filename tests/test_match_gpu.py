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
