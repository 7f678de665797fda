"""Oracle: CPU restatement of MapPoint::ComputeDistinctiveDescriptors.  TEST INFRASTRUCTURE ONLY.

Follows src/MapPoint.cc:331-400: for one map point with N observed 256-d descriptors, ``Distances[i][j] =
Matcher::DescriptorDistance(d_i, d_j)`` (``(a-b).norm()`` in fp32, src/Matcher.cc:1893-1900, zero diagonal), each row
sorted, ``median = sorted[int(0.5 * (N - 1))]``, and the representative descriptor is the FIRST row whose median is
strictly smaller than every earlier one (``median < BestMedian``, :388-392).

Pin: ``distinctive_index`` (vectorised) is checked against ``distinctive_index_literal`` (the loops of the reference
transcribed line by line) in tests/test_oracle_pins.py; Eigen's summation order inside ``norm()`` is unpinned, so medians
are compared with a tolerance and indices exactly away from fp32 ties.
"""
from __future__ import annotations

import numpy as np


def distance_matrix(D: np.ndarray) -> np.ndarray:
    """Distances[i][j] = ||d_i - d_j||_2 in fp32 (difference first, like Eigen's (a-b).norm()), zero diagonal."""
    D = np.asarray(D, np.float32)
    diff = D[:, None, :] - D[None, :, :]
    M = np.sqrt(np.sum(diff * diff, axis=2, dtype=np.float32)).astype(np.float32)
    np.fill_diagonal(M, np.float32(0))
    return M


def distinctive_index(D: np.ndarray):
    """Returns (best_index, best_median) for one map point's descriptors D [N][256]; (-1, fmax) when N == 0."""
    N = int(np.asarray(D).shape[0])
    if N == 0:
        return -1, np.finfo(np.float32).max
    M = np.sort(distance_matrix(D), axis=1)
    med = M[:, int(0.5 * (N - 1))]
    best = int(np.argmin(med))          # argmin returns the first minimum == the reference's strict '<' scan
    return best, np.float32(med[best])


def distinctive_index_literal(D: np.ndarray):
    """The reference's loops, one to one (MapPoint.cc:368-394)."""
    D = np.asarray(D, np.float32)
    N = D.shape[0]
    if N == 0:
        return -1, np.finfo(np.float32).max
    dist = np.zeros((N, N), np.float32)
    for i in range(N):
        dist[i, i] = 0
        for j in range(i + 1, N):
            d = (D[i] - D[j]).astype(np.float32)
            dij = np.float32(np.sqrt(np.sum(d * d, dtype=np.float32)))
            dist[i, j] = dij
            dist[j, i] = dij
    best_median = np.finfo(np.float32).max
    best_idx = 0
    for i in range(N):
        v = sorted(dist[i].tolist())
        median = np.float32(v[int(0.5 * (N - 1))])
        if median < best_median:
            best_median = median
            best_idx = i
    return best_idx, best_median


def distinctive_batch(descriptors: np.ndarray, offsets: np.ndarray):
    """Ragged batch: map point p owns rows offsets[p]:offsets[p+1].  Returns (best_index int32[n], best_median f32[n]),
    indices relative to the point's first row."""
    n = len(offsets) - 1
    idx = np.full(n, -1, np.int32)
    med = np.full(n, np.finfo(np.float32).max, np.float32)
    for p in range(n):
        idx[p], med[p] = distinctive_index(descriptors[offsets[p]:offsets[p + 1]])
    return idx, med
