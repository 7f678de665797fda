"""Host-side mirror of ``KeyFrameDatabase`` (include/KeyFrameDatabase.h:59-69, src/KeyFrameDatabase.cc) over the GPU scan.

The 4096-d scan over every keyframe (src/KeyFrameDatabase.cc:85-96) runs on the device through ``hfb_kfdb_*``; the
covisibility accumulation (:111-137), the sort (:139) and the candidate pick (:146-166) walk the caller's keyframe
graph and stay on the host, exactly as SURVEY.md section 2 row 8 scopes it.  The reference's infinite loop on a bad
keyframe (:149-150) is not reproduced: bad keyframes are skipped.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .lib import Context, _f32p, _i32p, _i64p, ptr


class KeyFrameDatabase:
    def __init__(self, ctx: Context, capacity: int = 65536, dim: int = 4096):
        self.ctx, self.dim, self.capacity = ctx, dim, capacity
        self.handle = C.c_void_p()
        ctx.check(ctx.lib.hfb_kfdb_create(ctx.handle, dim, capacity, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.hfb_kfdb_destroy(self.handle)
        self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.ctx.lib.hfb_kfdb_size(self.handle))

    # KeyFrameDatabase::add / erase / clear (src/KeyFrameDatabase.cc:31-52)
    def add(self, kf_id: int, global_descriptor: np.ndarray):
        self.add_many(np.array([kf_id], np.int64), np.asarray(global_descriptor, np.float32).reshape(1, self.dim))

    def add_many(self, ids: np.ndarray, descriptors: np.ndarray):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        d = np.ascontiguousarray(descriptors, dtype=np.float32).reshape(len(ids), self.dim)
        self.ctx.check(self.ctx.lib.hfb_kfdb_add(self.handle, ptr(ids, _i64p), ptr(d, _f32p), len(ids)))

    def add_tagged(self, ids: np.ndarray, map_ids: np.ndarray, descriptors: np.ndarray):
        """add with KeyFrame::GetMap() recorded (for clear_map)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        maps = np.ascontiguousarray(map_ids, dtype=np.int64)
        d = np.ascontiguousarray(descriptors, dtype=np.float32).reshape(len(ids), self.dim)
        self.ctx.check(self.ctx.lib.hfb_kfdb_add_tagged(self.handle, ptr(ids, _i64p), ptr(maps, _i64p), ptr(d, _f32p),
                                                        len(ids)))

    def clear_map(self, map_id: int):
        """KeyFrameDatabase::clearMap (src/KeyFrameDatabase.cc:54-68)."""
        self.ctx.check(self.ctx.lib.hfb_kfdb_clear_map(self.handle, int(map_id)))

    def erase(self, kf_id: int):
        self.ctx.check(self.ctx.lib.hfb_kfdb_erase(self.handle, int(kf_id)))

    def clear(self):
        self.ctx.check(self.ctx.lib.hfb_kfdb_clear(self.handle))

    # ------------------------------------------------------------------------------------------------ device scan
    def query(self, q: np.ndarray, rel: float = 0.8, floor: float = 0.0) -> Tuple[np.ndarray, np.ndarray, float]:
        """Returns (candidate ids ascending, their scores, best score): score > max(floor, rel*best), strict."""
        qq = np.ascontiguousarray(q, dtype=np.float32).reshape(self.dim)
        cap = max(len(self), 1)
        buf = getattr(self, "_query_out", None)
        if buf is None or len(buf[0]) < cap:       # output arrays kept across calls (a fresh 600 KB block per query costs more
            buf = self._query_out = (np.zeros(max(cap, 1024), np.int64), np.zeros(max(cap, 1024), np.float32))   # than the call)
        ids, sc = buf
        n, best = C.c_int32(), C.c_float()
        self.ctx.check(self.ctx.lib.hfb_kfdb_query(self.handle, ptr(qq, _f32p), rel, floor, ptr(ids, _i64p),
                                                   ptr(sc, _f32p), len(ids), C.byref(n), C.byref(best)))
        return ids[:n.value].copy(), sc[:n.value].copy(), float(best.value)

    def query_batch(self, Q: np.ndarray, rel: float = 0.8, floor: float = 0.0, cap: int = 256):
        """Several queries as one pass over the rows (tensor-core selection + exact re-scoring): list of
        (candidate ids ascending, scores, best) per query, equal to calling query() for each."""
        qq = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, self.dim)
        nq = qq.shape[0]
        ids = np.zeros((nq, cap), np.int64)
        sc = np.zeros((nq, cap), np.float32)
        n = np.zeros(nq, np.int32)
        best = np.zeros(nq, np.float32)
        self.ctx.check(self.ctx.lib.hfb_kfdb_query_batch(self.handle, ptr(qq, _f32p), nq, rel, floor, cap, ptr(ids, _i64p),
                                                         ptr(sc, _f32p), ptr(n, _i32p), ptr(best, _f32p)))
        return [(ids[i, :n[i]].copy(), sc[i, :n[i]].copy(), float(best[i])) for i in range(nq)]

    def scores_of(self, ids: np.ndarray) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.zeros(len(ids), np.float32)
        self.ctx.check(self.ctx.lib.hfb_kfdb_scores_of(self.handle, ptr(ids, _i64p), len(ids), ptr(out, _f32p)))
        return out

    def query_shard(self, q: np.ndarray, rel: float = 0.8, floor: float = 0.0, k: int = 64) -> bytes:
        """Fixed-size record (16 + 16k bytes) of this shard for one all-gather; see merge_shard_records."""
        qq = np.ascontiguousarray(q, dtype=np.float32).reshape(self.dim)
        buf = C.create_string_buffer(16 + 16 * k)
        self.ctx.check(self.ctx.lib.hfb_kfdb_query_shard(self.handle, ptr(qq, _f32p), rel, floor, k, buf))
        return buf.raw

    # ------------------------------------------------------------------------------------------------ sharded, on device
    def shard_setup(self, rank: int, world: int, k: int = 64) -> bytes:
        """Allocate this rank's inbox; returns the 64-byte IPC handle the other ranks connect with."""
        h = C.create_string_buffer(64)
        self.ctx.check(self.ctx.lib.hfb_kfdb_shard_setup(self.handle, rank, world, k, h))
        self._shard = (rank, world, k)
        return h.raw

    def connect_shards(self, dist, rank: int, world: int, k: int = 64):
        """Multi-process start-up: one all_gather of the IPC handles (torch.distributed, any backend), then every rank maps
        its peers' inboxes."""
        import torch
        mine = self.shard_setup(rank, world, k)
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        blob = b"".join(bytes(o.cpu().numpy()) for o in out)
        self.ctx.check(self.ctx.lib.hfb_kfdb_shard_connect(self.handle, blob))

    @staticmethod
    def connect_shards_local(shards: Sequence["KeyFrameDatabase"], k: int = 64):
        """Same-process variant (several shard objects, e.g. one per context on one device)."""
        world = len(shards)
        for r, sh in enumerate(shards):
            sh.shard_setup(r, world, k)
        arr = (C.c_void_p * world)(*[sh.handle for sh in shards])
        for sh in shards:
            sh.ctx.check(sh.ctx.lib.hfb_kfdb_shard_connect_local(sh.handle, arr))

    def query_sharded_begin(self, q: np.ndarray, rel: float = 0.8, floor: float = 0.0):
        qq = np.ascontiguousarray(q, dtype=np.float32).reshape(self.dim)
        self.ctx.check(self.ctx.lib.hfb_kfdb_query_sharded_begin(self.handle, ptr(qq, _f32p), rel, floor))

    def query_sharded_end(self):
        _, world, k = self._shard
        cap = world * k
        ids, sc = np.zeros(cap, np.int64), np.zeros(cap, np.float32)
        n, best, ov = C.c_int32(), C.c_float(), C.c_int32()
        self.ctx.check(self.ctx.lib.hfb_kfdb_query_sharded_end(self.handle, ptr(ids, _i64p), ptr(sc, _f32p), cap, C.byref(n),
                                                               C.byref(best), C.byref(ov)))
        return ids[:n.value].copy(), sc[:n.value].copy(), float(best.value), bool(ov.value)

    def query_sharded(self, q: np.ndarray, rel: float = 0.8, floor: float = 0.0):
        """Collective query over the sharded database: (global candidate ids ascending, scores, global best, overflow)."""
        self.query_sharded_begin(q, rel, floor)
        return self.query_sharded_end()

    # ------------------------------------------------------------------------------------------------ host logic
    def _accumulate(self, cand: Sequence[int], sc: Dict[int, float], covisibles: Callable[[int, int], Iterable[int]]):
        """src/KeyFrameDatabase.cc:111-137.  Neighbours that are not candidates themselves still count if the
        database scored them this query (every keyframe in the database is scored, :86-96)."""
        nbr_lists = [list(covisibles(kf, 10)) for kf in cand]
        need = sorted({nb for lst in nbr_lists for nb in lst if nb not in sc})
        if need:
            vals = self.scores_of(np.array(need, np.int64))
            for nb, v in zip(need, vals):
                if v >= 0:
                    sc[nb] = float(v)
        out = []
        for kf, nbrs in zip(cand, nbr_lists):
            best_score = np.float32(sc[kf])
            acc = np.float32(best_score)
            best_kf = kf
            for nb in nbrs:
                if nb not in sc:
                    continue
                s = np.float32(sc[nb])
                acc = np.float32(acc + s)
                if s > best_score:
                    best_kf, best_score = nb, s
            out.append((float(acc), best_kf))
        return out

    def detect_n_best_candidates(self, q: np.ndarray, query_map: int, map_of: Dict[int, int],
                                 covisibles: Callable[[int, int], Iterable[int]], n_candidates: int = 3,
                                 bad: Optional[set] = None) -> Tuple[List[int], List[int]]:
        """KeyFrameDatabase::DetectNBestCandidates (src/KeyFrameDatabase.cc:75-167) -> (loop ids, merge ids)."""
        ids, scores, _ = self.query(q, 0.8, 0.0)
        sc = {int(i): float(s) for i, s in zip(ids, scores)}
        cand = [int(i) for i in ids]
        acc = self._accumulate(cand, sc, covisibles)
        acc.sort(key=lambda t: -t[0])
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
        return loop, merge

    def detect_relocalization_candidates(self, q: np.ndarray, query_map: int, map_of: Dict[int, int],
                                         covisibles: Callable[[int, int], Iterable[int]]) -> List[int]:
        """KeyFrameDatabase::DetectRelocalizationCandidates (src/KeyFrameDatabase.cc:170-256)."""
        ids, scores, _ = self.query(q, 0.8, 0.5)
        sc = {int(i): float(s) for i, s in zip(ids, scores)}
        cand = [int(i) for i in ids]
        acc = self._accumulate(cand, sc, covisibles)
        best_acc = max([a for a, _ in acc], default=0.0)
        acc.sort(key=lambda t: -t[0])
        min_retain = np.float32(0.75) * np.float32(best_acc)
        keep, seen = [], set()
        for a, kf in acc:
            if np.float32(a) > min_retain and map_of[kf] == query_map and kf not in seen:
                keep.append(kf)
                seen.add(kf)
        return keep


def parse_shard_record(rec: bytes):
    best, count, overflow, _ = struct.unpack_from("<fiii", rec, 0)
    ids, scores = [], []
    for i in range(count):
        s, _, kid = struct.unpack_from("<fiq", rec, 16 + 16 * i)
        ids.append(kid)
        scores.append(s)
    return best, np.array(ids, np.int64), np.array(scores, np.float32), bool(overflow)


def merge_shard_records(records: Sequence[bytes], rel: float = 0.8, floor: float = 0.0):
    """After the all-gather: global best = max of the shard bests; every shard listed the rows above
    rel * ITS best (a superset of its share, local best <= global best), so filtering the union by the global bar gives
    exactly the unsharded candidate set unless a shard overflowed its k slots.  Returns (ids ascending, scores, best,
    overflow)."""
    parsed = [parse_shard_record(r) for r in records]
    best = np.float32(max([p[0] for p in parsed], default=0.0))
    thr = max(np.float32(floor), np.float32(best * np.float32(rel)))
    ids = np.concatenate([p[1] for p in parsed]) if parsed else np.zeros(0, np.int64)
    sc = np.concatenate([p[2] for p in parsed]) if parsed else np.zeros(0, np.float32)
    keep = sc > thr
    ids, sc = ids[keep], sc[keep]
    order = np.argsort(ids, kind="stable")
    # an overflowing shard only matters if its k-th entry is still above the global bar
    overflow = bool(any(p[3] and len(p[2]) and p[2].min() > thr for p in parsed))
    return ids[order], sc[order], float(best), overflow
