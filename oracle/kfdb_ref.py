"""Oracle: CPU restatement of the 4096-d keyframe-database place-recognition query.  TEST INFRASTRUCTURE ONLY.

Follows src/KeyFrameDatabase.cc:
* ``scores``                              :85-96   score = max(0, 1 - ||q - d||_2) over EVERY keyframe (fp32)
* ``candidate_set``                       :98-104  score > 0.8 * best (strict)
* ``detect_n_best_candidates``            :75-167  covisibility accumulation (10 best covisibles), sort by acc score
                                                    descending, first N loop / N merge candidates
* ``detect_relocalization_candidates``    :170-256 floor max(0.5, 0.8*best); keep acc > 0.75*bestAcc in the query map

The reference iterates an ``unordered_set<KeyFrame*>`` (hash order) and ``list::sort`` only on the accumulated score,
so candidate order among equal scores is unspecified: compare as SETS.  Here iteration is in
ascending keyframe id and the sort is stable.  The reference's ``if(pKFi->isBad()) continue;`` without advancing the
iterator (:149-150) spins forever on a bad keyframe; it is NOT reproduced (bad keyframes are skipped).

PARITY UNPINNED beyond the arithmetic definition: Eigen's fp32 ``norm()`` summation order is not specified; scores
here are computed in float64 and rounded to fp32.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def scores(q: np.ndarray, db: np.ndarray) -> np.ndarray:
    """max(0, 1 - ||q - d_i||) for every row of db.  q [4096], db [N,4096] -> f32[N]."""
    q64 = q.astype(np.float64)[None, :]
    nrm = np.empty(db.shape[0], np.float32)
    for r0 in range(0, db.shape[0], 4096):          # chunked: a 50 k-row database would need 1.6 GB in one piece
        diff = db[r0:r0 + 4096].astype(np.float64) - q64
        nrm[r0:r0 + 4096] = np.sqrt((diff * diff).sum(axis=1)).astype(np.float32)
    return np.maximum(np.float32(0), np.float32(1) - nrm).astype(np.float32)


def candidate_set(sc: np.ndarray, rel: float = 0.8, floor: Optional[float] = None) -> Tuple[np.ndarray, float]:
    best = np.float32(sc.max()) if sc.size else np.float32(0)
    best = max(best, np.float32(0))
    min_score = np.float32(best * np.float32(rel))
    if floor is not None:
        min_score = max(np.float32(floor), min_score)
    return np.flatnonzero(sc > min_score), float(best)


def _accumulate(cand: Sequence[int], sc_by_id: Dict[int, float], covis: Callable[[int, int], Iterable[int]]):
    """KeyFrameDatabase.cc:111-137.  Returns list of (accScore, bestKF id) in candidate iteration order."""
    out = []
    for kf in cand:
        best_score = np.float32(sc_by_id[kf])
        acc = np.float32(best_score)
        best_kf = kf
        for nb in covis(kf, 10):
            if nb not in sc_by_id:          # mnPlaceRecognitionQuery != query id  (not scored this query)
                continue
            s = np.float32(sc_by_id[nb])
            acc = np.float32(acc + s)
            if s > best_score:
                best_kf, best_score = nb, s
        out.append((float(acc), best_kf))
    return out


def detect_n_best_candidates(q: np.ndarray, ids: np.ndarray, db: np.ndarray, map_of: Dict[int, int], query_map: int,
                             covis: Callable[[int, int], Iterable[int]], n_candidates: int = 3,
                             bad: Optional[set] = None):
    """Returns (loop ids, merge ids, candidate id set)."""
    sc = scores(q, db)
    sel, _ = candidate_set(sc, 0.8)
    sc_by_id = {int(i): float(s) for i, s in zip(ids, sc)}
    cand = sorted(int(ids[i]) for i in sel)
    acc = _accumulate(cand, sc_by_id, covis)
    acc.sort(key=lambda t: -t[0])            # stable, descending accScore  (list::sort(compFirst))
    loop, merge, seen = [], [], set()
    for _, kf in acc:
        if len(loop) >= n_candidates and len(merge) >= n_candidates:
            break
        if bad and kf in bad:
            continue
        if kf not in seen:
            if map_of[kf] == query_map and len(loop) < n_candidates:
                loop.append(kf)
            elif map_of[kf] != query_map and len(merge) < n_candidates:
                merge.append(kf)
            seen.add(kf)
    return loop, merge, set(cand)


def detect_relocalization_candidates(q: np.ndarray, ids: np.ndarray, db: np.ndarray, map_of: Dict[int, int],
                                     query_map: int, covis: Callable[[int, int], Iterable[int]]) -> List[int]:
    sc = scores(q, db)
    sel, _ = candidate_set(sc, 0.8, floor=0.5)
    sc_by_id = {int(i): float(s) for i, s in zip(ids, sc)}
    cand = sorted(int(ids[i]) for i in sel)
    acc = _accumulate(cand, sc_by_id, covis)
    best_acc = max([a for a, _ in acc], default=0.0)
    acc.sort(key=lambda t: -t[0])
    keep, seen = [], set()
    min_retain = np.float32(0.75) * np.float32(best_acc)
    for a, kf in acc:
        if np.float32(a) > min_retain:
            if map_of[kf] != query_map:
                continue
            if kf not in seen:
                keep.append(kf)
                seen.add(kf)
    return keep


def topk_l2(q: np.ndarray, db: np.ndarray, k: int = 3) -> np.ndarray:
    """Examples/Utility/test_match_global_feats.cc:64-83: indices of the k smallest L2 distances (ties -> lower id)."""
    diff = db.astype(np.float64) - q.astype(np.float64)[None, :]
    d = np.sqrt((diff * diff).sum(axis=1)).astype(np.float32)
    return np.lexsort((np.arange(d.size), d))[:k]
