"""HF-Net architecture table, flat weight-blob format and a seeded synthetic weight generator.

The reference builds its network from ``HF-Net.onnx`` (src/Extractors/HFNetRTModel.cc:208-254); neither the
ONNX file nor the checkpoint is in the repository (README.md:62), so this module defines our own flat blob
(``HFB2WTS1``) that the C-ABI ``hfb_load_weights`` consumes.  A converter from ONNX initialisers only has to
emit tensors in the order of :func:`tensor_specs` with BatchNorm folded (see :func:`fold_bn`).

Architecture (hfnet/models/hf_net.py:13-52 ``MOBILENET_DEF``, depth multiplier 0.75 inferred from the 96-channel
``layer_7`` endpoint, src/Extractors/BaseModel.cc:70; channel rounding hfnet/models/backbones/utils/mobilenet.py:62-69,
expansion hfnet/models/backbones/utils/conv_blocks.py:158-159):

    conv1 3x3/2 1->24 | 17 inverted-residual blocks | local head (hf_net.py:55-96) | NetVLAD + FC (layers.py:57-109)

All tensors are fp32, BatchNorm (eps 1e-3, slim default) already folded into a per-output-channel scale (merged
into the weights) and bias.  Weight layouts are TensorFlow's HWIO flattened to ``[K][N]`` row-major with
``k = tap * Cin + ci`` (tap = 3*ky + kx), depthwise kernels ``[9][C]``.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

MAGIC = b"HFB2WTS1"
VERSION = 1
BN_EPS = 1e-3            # slim.batch_norm default epsilon (no override in hf_net.py / mobilenet.py)
DESC_DIM = 256           # hf_net.py:175 'descriptor_dim'
DET_GRID = 8             # hf_net.py:176 'detector_grid'
GLOBAL_DIM = 4096        # src/Extractors/HFNetRTModel.cc:201
LOCAL_ENDPOINT = 7       # 'layer_7'  (hf_net.py:163)
GLOBAL_ENDPOINT = 18     # 'layer_18' (hf_net.py:162)


def make_divisible(v: float, divisor: int, min_value: int | None = None) -> int:
    """hfnet/models/backbones/utils/mobilenet.py:62-69."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


@dataclass(frozen=True)
class Block:
    """One ``expanded_conv`` (conv_blocks.py:162-312): 1x1 expand -> 3x3 depthwise -> 1x1 project."""
    layer: int       # 'layer_N' endpoint index (conv1 is layer_1)
    cin: int
    cexp: int        # == cin  => no expand conv (conv_blocks.py:263-271)
    cout: int
    stride: int

    @property
    def has_expand(self) -> bool:
        return self.cexp > self.cin

    @property
    def residual(self) -> bool:
        # conv_blocks.py:302-311: stride 1 and depth matches
        return self.stride == 1 and self.cin == self.cout


# (stride, nominal num_outputs) for layers 2..18, hf_net.py:30-49
_NOMINAL = [(1, 16), (2, 24), (1, 24), (2, 32), (1, 64), (1, 128), (2, 64), (1, 64), (1, 64), (1, 64),
            (1, 96), (1, 96), (1, 96), (2, 160), (1, 160), (1, 160), (1, 320)]


def architecture(depth_multiplier: float = 0.75) -> Tuple[int, List[Block]]:
    """Returns (conv1 output channels, blocks for layer_2..layer_18)."""
    c1 = make_divisible(32 * depth_multiplier, 8, 8)
    blocks: List[Block] = []
    cin = c1
    for i, (stride, nominal) in enumerate(_NOMINAL):
        layer = i + 2
        cout = make_divisible(nominal * depth_multiplier, 8, 8)
        if layer == 2:
            cexp = make_divisible(cin * 1, 1)      # expand_input_by_factor(1, divisible_by=1), hf_net.py:32-34
        else:
            cexp = make_divisible(cin * 6, 8)      # expand_input_by_factor(6), hf_net.py:22
        blocks.append(Block(layer, cin, cexp, cout, stride))
        cin = cout
    return c1, blocks


def tensor_specs(n_clusters: int = 32, depth_multiplier: float = 0.75) -> List[Tuple[str, Tuple[int, ...]]]:
    """Canonical (name, shape) order of the blob payload."""
    c1, blocks = architecture(depth_multiplier)
    specs: List[Tuple[str, Tuple[int, ...]]] = [("conv1.w", (9, c1)), ("conv1.b", (c1,))]
    for b in blocks:
        p = f"l{b.layer}"
        if b.has_expand:
            specs += [(f"{p}.expand.w", (b.cin, b.cexp)), (f"{p}.expand.b", (b.cexp,))]
        specs += [(f"{p}.dw.w", (9, b.cexp)), (f"{p}.dw.b", (b.cexp,)),
                  (f"{p}.project.w", (b.cexp, b.cout)), (f"{p}.project.b", (b.cout,))]
    c_local = blocks[LOCAL_ENDPOINT - 2].cout
    c_global = blocks[GLOBAL_ENDPOINT - 2].cout
    specs += [("desc.conv1.w", (9 * c_local, DESC_DIM)), ("desc.conv1.b", (DESC_DIM,)),
              ("desc.conv2.w", (DESC_DIM, DESC_DIM)), ("desc.conv2.b", (DESC_DIM,)),
              ("det.conv1.w", (9 * c_local, 128)), ("det.conv1.b", (128,)),
              ("det.conv2.w", (128, DET_GRID * DET_GRID + 1)), ("det.conv2.b", (DET_GRID * DET_GRID + 1,)),
              ("vlad.memberships.w", (c_global, n_clusters)), ("vlad.memberships.b", (n_clusters,)),
              ("vlad.clusters", (n_clusters, c_global)),
              ("fc.w", (c_global * n_clusters, GLOBAL_DIM)), ("fc.b", (GLOBAL_DIM,))]
    return specs


def fold_bn(w: np.ndarray, gamma, beta, mean, var, eps: float = BN_EPS):
    """Fold inference BatchNorm into the preceding bias-free conv: y = (conv(x,w) - mean) * gamma/sqrt(var+eps) + beta.

    ``w`` has the output channel as its last axis.  Done in float64, returned as float32."""
    scale = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps)
    wf = np.asarray(w, np.float64) * scale
    bf = np.asarray(beta, np.float64) - np.asarray(mean, np.float64) * scale
    return wf.astype(np.float32), bf.astype(np.float32)


def pack(tensors: Dict[str, np.ndarray], n_clusters: int = 32, depth_multiplier: float = 0.75) -> bytes:
    """Serialise to the HFB2WTS1 blob: magic | u32 version | u32 n_clusters | f32 depth_multiplier | u32 n_tensors |
    u64 payload_floats | fp32 payload in :func:`tensor_specs` order."""
    specs = tensor_specs(n_clusters, depth_multiplier)
    chunks = []
    total = 0
    for name, shape in specs:
        t = np.ascontiguousarray(tensors[name], dtype=np.float32)
        if t.shape != tuple(shape):
            raise ValueError(f"{name}: expected {shape}, got {t.shape}")
        chunks.append(t.tobytes())
        total += t.size
    header = MAGIC + struct.pack("<IIfIQ", VERSION, n_clusters, depth_multiplier, len(specs), total)
    return header + b"".join(chunks)


HEADER_BYTES = len(MAGIC) + struct.calcsize("<IIfIQ")


def unpack(blob: bytes) -> Tuple[Dict[str, np.ndarray], int, float]:
    if blob[:8] != MAGIC:
        raise ValueError("bad magic")
    version, n_clusters, dm, n_tensors, total = struct.unpack_from("<IIfIQ", blob, 8)
    if version != VERSION:
        raise ValueError("bad version")
    specs = tensor_specs(n_clusters, dm)
    if n_tensors != len(specs):
        raise ValueError("tensor count mismatch")
    payload = np.frombuffer(blob, dtype=np.float32, offset=HEADER_BYTES)
    if payload.size != total:
        raise ValueError("payload size mismatch")
    out, off = {}, 0
    for name, shape in specs:
        n = int(np.prod(shape))
        out[name] = payload[off:off + n].reshape(shape)
        off += n
    if off != total:
        raise ValueError("payload size mismatch")
    return out, n_clusters, dm


def synthetic(seed: int = 0, n_clusters: int = 32, depth_multiplier: float = 0.75) -> Dict[str, np.ndarray]:
    """Seeded random-init weights of the real architecture: He-normal kernels, random BatchNorm
    statistics folded, detector logits scaled so that the softmax is peaky and a few thousand pixels of a 752x480
    frame pass threshold 0.01 after NMS."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}

    def bn(c):
        return (rng.uniform(0.5, 1.5, c), rng.normal(0, 0.1, c), rng.normal(0, 0.1, c), rng.uniform(0.5, 1.5, c))

    for name, shape in tensor_specs(n_clusters, depth_multiplier):
        if not name.endswith(".w"):
            continue
        base = name[:-2]
        if ".dw" in name:
            fan_in = 9
        else:
            fan_in = shape[0]
        w = rng.normal(0, np.sqrt(2.0 / fan_in), shape)
        if base in ("desc.conv2", "det.conv2", "fc"):
            # plain bias, no normaliser (hf_net.py:66-72, layers.py:99-107)
            if base == "det.conv2":
                w *= 0.45                     # moderately peaky softmax
                b = rng.normal(0, 0.5, shape[-1])
                b[-1] += 1.0                  # dustbin
            elif base == "fc":
                w = rng.normal(0, np.sqrt(1.0 / fan_in), shape)
                b = rng.normal(0, 0.01, shape[-1])
            else:
                w = rng.normal(0, np.sqrt(1.0 / fan_in), shape)
                b = rng.normal(0, 0.1, shape[-1])
            out[name] = w.astype(np.float32)
            out[base + ".b"] = b.astype(np.float32)
        else:
            g, be, mu, var = bn(shape[-1])
            if base.endswith(".project") or base == "vlad.memberships":
                pass                           # linear (BN only)
            wf, bf = fold_bn(w, g, be, mu, var)
            out[name] = wf
            out[base + ".b"] = bf
    c_global = architecture(depth_multiplier)[1][GLOBAL_ENDPOINT - 2].cout
    out["vlad.clusters"] = rng.normal(0, 1.0, (n_clusters, c_global)).astype(np.float32)
    return out


def synthetic_blob(seed: int = 0, n_clusters: int = 32) -> bytes:
    return pack(synthetic(seed, n_clusters), n_clusters)


def synthetic_image(height: int = 480, width: int = 752, seed: int = 1, n_corners: int = 200) -> np.ndarray:
    """u8 test frame: blurred noise + random bright corner blobs.  numpy only."""
    rng = np.random.default_rng(seed)
    img = rng.normal(110.0, 40.0, (height, width))
    k = np.array([1, 4, 6, 4, 1], np.float64) / 16.0
    for _ in range(2):
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, img)
        img = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, img)
    ys = rng.integers(4, height - 4, n_corners)
    xs = rng.integers(4, width - 4, n_corners)
    for y, x in zip(ys, xs):
        img[y:y + 3, x:x + 3] += 120.0
        img[y - 3:y, x - 3:x] -= 60.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
