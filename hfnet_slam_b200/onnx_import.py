"""ONNX initialisers -> the HFB2WTS1 weight blob, without an ONNX dependency.

The reference ships its network as ``HF-Net.onnx`` (hfnet/README.md:28-43, loaded by src/Extractors/HFNetRTModel.cc:208-227);
neither that file nor the ``onnx`` package exists in this image, so this module reads the protobuf wire format directly
(ModelProto.graph = field 7; GraphProto.node = 1, .initializer = 5; NodeProto.input = 1, .op_type = 4; TensorProto.dims = 1,
.data_type = 2, .float_data = 4, .name = 8, .raw_data = 9) and assigns tensors BY SHAPE AND GRAPH ORDER, not by name:

* every ``Conv`` node contributes (weight initialiser, optional bias initialiser); convolutions that share a weight shape
  are sequentially dependent in HF-Net (e.g. layer_9 / 10 / 11), so their order in the topologically sorted node list is
  their order in the network;
* ``BatchNormalization`` nodes that follow a bias-free Conv are folded (weights.fold_bn); exporters that already folded
  them (tf2onnx does) simply provide the bias;
* the NetVLAD centroids are the float initialiser with n_clusters * c_global elements that is not a Conv weight, the
  dimensionality reduction is the 2-D initialiser with 4096 on one side (MatMul [K, 4096] or Gemm [4096, K]) and its
  bias the 4096-vector.

ONNX convolution weights are OIHW; the blob keeps TensorFlow's HWIO flattened to [kh*kw*cin, cout] (depthwise:
[9, channels]).  Without the real file the naming and node layout of ``HF-Net.onnx`` are unverified: ``convert`` checks
every shape and fails loudly on the first mismatch instead of guessing.  ``write_model`` emits the same subset of ONNX
and exists for the round-trip test (tests/test_host_logic_cpu.py) and as an executable statement of the convention.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import weights as W

_FLOAT, _FLOAT16 = 1, 10


# ---------------------------------------------------------------------------------------------- protobuf wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yields (field number, wire type, value) of one message; value = int (varint / fixed) or bytes (length-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _tensor(buf: bytes) -> Tuple[str, Optional[np.ndarray]]:
    dims: List[int] = []
    dtype, name, raw, floats = 0, "", None, []
    for num, wt, v in _fields(buf):
        if num == 1:
            if wt == 2:                      # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(d)
            else:
                dims.append(v)
        elif num == 2:
            dtype = v
        elif num == 4:
            floats.append(np.frombuffer(v, "<f4") if wt == 2 else np.array([struct.unpack("<f", struct.pack("<I", v))[0]], "<f4"))
        elif num == 8:
            name = v.decode()
        elif num == 9:
            raw = v
    if dtype == _FLOAT:
        a = np.frombuffer(raw, "<f4") if raw is not None else (np.concatenate(floats) if floats else np.zeros(0, "<f4"))
    elif dtype == _FLOAT16 and raw is not None:
        a = np.frombuffer(raw, "<f2").astype(np.float32)
    else:
        return name, None                    # integer constants (shapes, axes): not weights
    return name, np.array(a, np.float32).reshape(dims)


def read_model(data) -> Tuple[List[Tuple[str, List[str]]], Dict[str, np.ndarray]]:
    """(nodes, initialisers): nodes = [(op_type, input names)] in file (topological) order; initialisers = float tensors."""
    buf = data if isinstance(data, (bytes, bytearray)) else open(data, "rb").read()
    graph = None
    for num, wt, v in _fields(bytes(buf)):
        if num == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError("no GraphProto (field 7) in the model")
    nodes, inits = [], {}
    for num, wt, v in _fields(graph):
        if num == 1 and wt == 2:
            op, ins = "", []
            for n2, w2, v2 in _fields(v):
                if n2 == 1:
                    ins.append(v2.decode())
                elif n2 == 4:
                    op = v2.decode()
            nodes.append((op, ins))
        elif num == 5 and wt == 2:
            name, arr = _tensor(v)
            if arr is not None:
                inits[name] = arr
    return nodes, inits


def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(num: int, payload: bytes) -> bytes:
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def write_model(path, convs: List[Tuple[str, np.ndarray, Optional[np.ndarray]]], extra: Dict[str, np.ndarray]) -> None:
    """Minimal ONNX file: one Conv node per (name, OIHW weight, bias) in the given order + extra float initialisers."""
    graph = bytearray()
    inits = bytearray()

    def tensor(name, a):
        a = np.ascontiguousarray(a, "<f4")
        t = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in a.shape)
        t += _enc_varint((2 << 3) | 0) + _enc_varint(_FLOAT) + _ld(8, name.encode()) + _ld(9, a.tobytes())
        return _ld(5, t)

    for name, w, b in convs:
        ins = ["x", name + "/W"] + ([name + "/B"] if b is not None else [])
        node = b"".join(_ld(1, s.encode()) for s in ins) + _ld(2, (name + "/Y").encode()) + _ld(4, b"Conv")
        graph += _ld(1, node)
        inits += tensor(name + "/W", w)
        if b is not None:
            inits += tensor(name + "/B", b)
    for name, a in extra.items():
        inits += tensor(name, a)
    model = _enc_varint((1 << 3) | 0) + _enc_varint(8) + _ld(7, bytes(graph) + bytes(inits))
    with open(path, "wb") as f:
        f.write(model)


# ---------------------------------------------------------------------------------------------- conversion
def _from_oihw(w: np.ndarray, depthwise: bool) -> np.ndarray:
    co, ci, kh, kw = w.shape
    if depthwise:
        return np.ascontiguousarray(w.transpose(2, 3, 0, 1).reshape(kh * kw, co))     # [9][C]
    return np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(kh * kw * ci, co))    # HWIO flattened


def to_oihw(w: np.ndarray, k: int, cin: int, depthwise: bool) -> np.ndarray:
    """Inverse of the above (used by the round-trip test and to document the convention)."""
    if depthwise:
        return np.ascontiguousarray(w.reshape(k, k, w.shape[1], 1).transpose(2, 3, 0, 1))
    return np.ascontiguousarray(w.reshape(k, k, cin, w.shape[1]).transpose(3, 2, 0, 1))


def expected_convs(n_clusters: int = 32, depth_multiplier: float = 0.75):
    """[(blob tensor base name, OIHW weight shape, depthwise?)] in network order."""
    c1, blocks = W.architecture(depth_multiplier)
    out = [("conv1", (c1, 1, 3, 3), False)]
    for b in blocks:
        p = f"l{b.layer}"
        if b.has_expand:
            out.append((p + ".expand", (b.cexp, b.cin, 1, 1), False))
        out.append((p + ".dw", (b.cexp, 1, 3, 3), True))
        out.append((p + ".project", (b.cout, b.cexp, 1, 1), False))
    c_local = blocks[W.LOCAL_ENDPOINT - 2].cout
    c_global = blocks[W.GLOBAL_ENDPOINT - 2].cout
    out += [("desc.conv1", (W.DESC_DIM, c_local, 3, 3), False), ("desc.conv2", (W.DESC_DIM, W.DESC_DIM, 1, 1), False),
            ("det.conv1", (128, c_local, 3, 3), False), ("det.conv2", (W.DET_GRID * W.DET_GRID + 1, 128, 1, 1), False),
            ("vlad.memberships", (n_clusters, c_global, 1, 1), False)]
    return out


def convert(data, n_clusters: int = 32, depth_multiplier: float = 0.75) -> Dict[str, np.ndarray]:
    """ONNX file / bytes -> dict of blob tensors (weights.pack() turns it into the HFB2WTS1 blob)."""
    nodes, inits = read_model(data)
    convs = []                                           # (weight, bias or None, BN params or None) in graph order
    used = set()
    for i, (op, ins) in enumerate(nodes):
        if op != "Conv" or len(ins) < 2 or ins[1] not in inits:
            continue
        w = inits[ins[1]]
        b = inits.get(ins[2]) if len(ins) > 2 else None
        bn = None
        if b is None and i + 1 < len(nodes) and nodes[i + 1][0] == "BatchNormalization" and len(nodes[i + 1][1]) >= 5:
            names = nodes[i + 1][1][1:5]
            if all(n in inits for n in names):
                bn = [inits[n] for n in names]          # scale, bias, mean, var (ONNX BatchNormalization input order)
                used.update(names)
        used.add(ins[1])
        if len(ins) > 2:
            used.add(ins[2])
        convs.append([w, b, bn, False])
    out: Dict[str, np.ndarray] = {}
    for base, shape, dw in expected_convs(n_clusters, depth_multiplier):
        hit = next((c for c in convs if not c[3] and tuple(c[0].shape) == tuple(shape)), None)
        if hit is None:
            have = sorted({tuple(c[0].shape) for c in convs if not c[3]})
            raise ValueError(f"{base}: no unused Conv with weight shape {shape} (OIHW); unused shapes: {have[:12]} ...")
        hit[3] = True
        w, b, bn, _ = hit
        w2 = _from_oihw(w, dw)
        if bn is not None:
            w2, b = W.fold_bn(w2, bn[0], bn[1], bn[2], bn[3])
        if b is None:
            raise ValueError(f"{base}: Conv without bias and without a following BatchNormalization")
        if b.shape != (shape[0],):
            raise ValueError(f"{base}: bias shape {b.shape}, expected ({shape[0]},)")
        out[base + ".w"], out[base + ".b"] = w2.astype(np.float32), np.asarray(b, np.float32)
    out["vlad.memberships.w"] = out["vlad.memberships.w"].reshape(-1, n_clusters)    # 1x1: [c_global][C]
    c_global = out["vlad.memberships.w"].shape[0]
    rest = {k: v for k, v in inits.items() if k not in used}
    cl = [v for v in rest.values() if v.size == n_clusters * c_global and v.ndim >= 2 and v.shape[-1] == c_global]
    fc = [v for v in rest.values() if v.ndim == 2 and W.GLOBAL_DIM in v.shape and v.size == c_global * n_clusters * W.GLOBAL_DIM]
    fb = [v for v in rest.values() if v.shape == (W.GLOBAL_DIM,)]
    if len(cl) != 1 or len(fc) != 1 or len(fb) != 1:
        raise ValueError(f"NetVLAD centroids / FC weight / FC bias candidates: {len(cl)} / {len(fc)} / {len(fb)} (expected 1 each)")
    out["vlad.clusters"] = cl[0].reshape(n_clusters, c_global).astype(np.float32)
    out["fc.w"] = (fc[0] if fc[0].shape[1] == W.GLOBAL_DIM else fc[0].T).astype(np.float32)
    out["fc.b"] = fb[0].astype(np.float32)
    for name, shp in W.tensor_specs(n_clusters, depth_multiplier):
        if name not in out or out[name].shape != tuple(shp):
            raise ValueError(f"{name}: converted shape {None if name not in out else out[name].shape}, blob expects {shp}")
    return out
