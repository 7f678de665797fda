"""Oracle: fp32 (or fp64) CPU restatement of the HF-Net inference graph.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference executes this graph through TensorRT 8.5 FP16 from an ONNX file that is not in the
repository (src/Extractors/HFNetRTModel.cc:130,208-254; README.md:62); no reference test pins its numerics.  This
file restates the graph definition op by op from the TF1-slim sources:

* image normalisation + crop      hfnet/models/hf_net.py:185-190, hfnet/models/utils/layers.py:6-7
* MobileNetV2 backbone            hfnet/models/hf_net.py:13-52, backbones/utils/conv_blocks.py:162-312
* local head                      hfnet/models/hf_net.py:55-96
* simple_nms (radius 4, 2 iters)  hfnet/models/utils/layers.py:10-32, hfnet/export_model.py:35-37
* NetVLAD + FC                    hfnet/models/utils/layers.py:57-109

TensorFlow 'SAME' padding is asymmetric on stride 2 (pad_before = total//2), which torch's ``padding=1`` is not,
so every convolution pads explicitly.  Convolutions run through torch CPU (MKL/oneDNN) in the requested dtype.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from hfnet_slam_b200.weights import (DET_GRID, GLOBAL_ENDPOINT, LOCAL_ENDPOINT, architecture)


def _same_pad(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _pad_same(x: torch.Tensor, k: int, s: int, value: float = 0.0) -> torch.Tensor:
    (t, b), (l, r) = _same_pad(x.shape[2], k, s), _same_pad(x.shape[3], k, s)
    return F.pad(x, (l, r, t, b), value=value)


def _conv(x, w_kn, b, k, s, cin, groups=1):
    """x NCHW; w_kn is the blob layout [k*k*cin(/groups)][cout] (HWIO flattened)."""
    cout = w_kn.shape[-1]
    if groups == 1:
        w = w_kn.reshape(k, k, cin, cout).permute(3, 2, 0, 1).contiguous()
    else:  # depthwise: [9][C] -> [C,1,3,3]
        w = w_kn.reshape(k, k, cout).permute(2, 0, 1).unsqueeze(1).contiguous()
    if k > 1:
        x = _pad_same(x, k, s)
    return F.conv2d(x, w, b, stride=s, groups=groups)


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def simple_nms(scores: torch.Tensor, radius: int = 4, iterations: int = 2) -> torch.Tensor:
    """hfnet/models/utils/layers.py:10-32.  scores [B,H,W].  Max-pool 'SAME' pads with -inf (TF semantics)."""
    size = 2 * radius + 1

    def max_pool(x):
        return F.max_pool2d(x[:, None], kernel_size=size, stride=1, padding=radius)[:, 0]

    zeros = torch.zeros_like(scores)
    max_mask = scores == max_pool(scores)
    for _ in range(iterations - 1):
        supp_mask = max_pool(max_mask.to(scores.dtype)) > 0
        supp_scores = torch.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == max_pool(supp_scores)
        max_mask = max_mask | (new_max_mask & ~supp_mask)
    return torch.where(max_mask, scores, zeros)


def forward(image_u8: np.ndarray, weights: Dict[str, np.ndarray], *, want_global: bool = True,
            dtype=torch.float32, nms_radius: int = 4, nms_iterations: int = 2,
            return_intermediates: bool = False) -> Dict[str, np.ndarray]:
    """image_u8: [H,W] or [B,H,W] uint8.  Returns numpy arrays:
    scores_dense [B,H8,W8], scores_dense_nms [B,H8,W8], local_descriptor_map [B,H8/8,W8/8,256] (unit rows),
    global_descriptor [B,4096] (if want_global), plus per-layer NHWC activations if return_intermediates."""
    img = np.asarray(image_u8)
    if img.ndim == 2:
        img = img[None]
    assert img.dtype == np.uint8
    w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
    n_clusters = w["vlad.clusters"].shape[0]
    c1, blocks = architecture()
    inter: Dict[str, np.ndarray] = {}

    def keep(name, t):
        if return_intermediates:
            inter[name] = t.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy()

    with torch.no_grad():
        x = torch.from_numpy(img.astype(np.float32)).to(dtype)
        x = (x - 128.0) / 128.0                                   # layers.py:6-7
        h8, w8 = (x.shape[1] // 8) * 8, (x.shape[2] // 8) * 8       # hf_net.py:188-190
        x = x[:, None, :h8, :w8]
        x = relu6(_conv(x, w["conv1.w"], w["conv1.b"], 3, 2, 1))
        keep("layer_1", x)
        local_feat = None
        for b in blocks:
            p = f"l{b.layer}"
            inp = x
            if b.has_expand:
                x = relu6(_conv(x, w[p + ".expand.w"], w[p + ".expand.b"], 1, 1, b.cin))
            x = relu6(_conv(x, w[p + ".dw.w"], w[p + ".dw.b"], 3, b.stride, b.cexp, groups=b.cexp))
            x = _conv(x, w[p + ".project.w"], w[p + ".project.b"], 1, 1, b.cexp)
            if b.residual:
                x = x + inp
            keep(f"layer_{b.layer}", x)
            if b.layer == LOCAL_ENDPOINT:
                local_feat = x
                if not want_global:
                    break
        c_local = local_feat.shape[1]
        # descriptor head, hf_net.py:74-80
        d = relu6(_conv(local_feat, w["desc.conv1.w"], w["desc.conv1.b"], 3, 1, c_local))
        keep("desc_conv1", d)
        d = _conv(d, w["desc.conv2.w"], w["desc.conv2.b"], 1, 1, d.shape[1])
        d = d.permute(0, 2, 3, 1)
        # tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))
        d = d * torch.rsqrt(torch.clamp((d * d).sum(-1, keepdim=True), min=1e-12))
        # detector head, hf_net.py:82-93
        l = relu6(_conv(local_feat, w["det.conv1.w"], w["det.conv1.b"], 3, 1, c_local))
        keep("det_conv1", l)
        l = _conv(l, w["det.conv2.w"], w["det.conv2.b"], 1, 1, l.shape[1])
        keep("det_logits", l)
        prob = torch.softmax(l.permute(0, 2, 3, 1), dim=-1)[..., :-1]           # strip dustbin
        B, Hc, Wc, _ = prob.shape
        g = DET_GRID
        # depth_to_space NHWC: out[8h+i, 8w+j] = in[h, w, 8i+j]
        prob = prob.reshape(B, Hc, Wc, g, g).permute(0, 1, 3, 2, 4).reshape(B, Hc * g, Wc * g)
        nms = simple_nms(prob, nms_radius, nms_iterations)
        out = {
            "scores_dense": prob.to(torch.float32).numpy(),
            "scores_dense_nms": nms.to(torch.float32).numpy(),
            "local_descriptor_map": d.contiguous().to(torch.float32).numpy(),
        }
        if want_global:
            f = x.permute(0, 2, 3, 1)                                            # [B,h,w,D]
            D = f.shape[-1]
            m = torch.matmul(f, w["vlad.memberships.w"]) + w["vlad.memberships.b"]
            m = torch.softmax(m, dim=-1)                                         # [B,h,w,C]
            cl = w["vlad.clusters"]                                              # [C,D]
            # layers.py:81-86: sum_hw m[hw,c] * (clusters[c,:] - x[hw,:])
            msum = m.sum(dim=(1, 2))                                             # [B,C]
            v = msum[:, :, None] * cl[None] - torch.einsum("bhwc,bhwd->bcd", m, f)
            if return_intermediates:
                inter["vlad_raw"] = v.to(torch.float32).numpy()
            # layers.py:88: l2_normalize(axis=1) on [B,C,D]  (over the CLUSTER axis, restated literally)
            v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
            v = v.reshape(B, n_clusters * D)
            v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
            # dimensionality_reduction, layers.py:95-109
            v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
            if return_intermediates:
                inter["vlad_norm"] = v.to(torch.float32).numpy()
            y = torch.matmul(v, w["fc.w"]) + w["fc.b"]
            y = y * torch.rsqrt(torch.clamp((y * y).sum(1, keepdim=True), min=1e-12))
            out["global_descriptor"] = y.to(torch.float32).numpy()
    out.update(inter)
    return out
