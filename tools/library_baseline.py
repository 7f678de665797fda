"""Library baseline of the HF-Net network forward on the SAME B200: the identical graph (hfnet/models/hf_net.py:13-104,
layers.py:57-109; BatchNorm folded) as cuDNN fp16 channels_last convolutions through PyTorch, replayed from a CUDA
graph.  This is the stand-in for the reference's TensorRT FP16 engine (src/Extractors/HFNetRTModel.cc:122-137), which cannot
be built in this image (no TensorRT): a vendor-library execution of the same layers, next to which the hand-written
kernels are read.  It covers the network only (dense score map after NMS, dense descriptor map, global descriptor) --
not the keypoint selection / sampling / matching that the product step also does.  Measurement tool: imported by
bench.py (``extra.library_baseline``), never by hfnet_slam_b200/."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from hfnet_slam_b200.weights import DET_GRID, LOCAL_ENDPOINT, architecture


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


class LibraryHFNet:
    def __init__(self, wd, dev, dtype=torch.float16):
        self.dev, self.dtype = dev, dtype
        self.c1, self.blocks = architecture()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        def conv_w(w_kn, k, cin, groups=1):
            cout = w_kn.shape[-1]
            if groups == 1:
                w = t(w_kn).reshape(k, k, cin, cout).permute(3, 2, 0, 1)
            else:
                w = t(w_kn).reshape(k, k, cout).permute(2, 0, 1).unsqueeze(1)
            return w.to(dtype).contiguous(memory_format=torch.channels_last)

        self.w = {}
        self.w["conv1"] = (conv_w(wd["conv1.w"], 3, 1), t(wd["conv1.b"]).to(dtype))
        for b in self.blocks:
            p = f"l{b.layer}"
            if b.has_expand:
                self.w[p + ".expand"] = (conv_w(wd[p + ".expand.w"], 1, b.cin), t(wd[p + ".expand.b"]).to(dtype))
            self.w[p + ".dw"] = (conv_w(wd[p + ".dw.w"], 3, b.cexp, groups=b.cexp), t(wd[p + ".dw.b"]).to(dtype))
            self.w[p + ".project"] = (conv_w(wd[p + ".project.w"], 1, b.cexp), t(wd[p + ".project.b"]).to(dtype))
        c_local = [b for b in self.blocks if b.layer == LOCAL_ENDPOINT][0].cout
        self.w["desc1"] = (conv_w(wd["desc.conv1.w"], 3, c_local), t(wd["desc.conv1.b"]).to(dtype))
        self.w["desc2"] = (conv_w(wd["desc.conv2.w"], 1, 256), t(wd["desc.conv2.b"]).to(dtype))
        self.w["det1"] = (conv_w(wd["det.conv1.w"], 3, c_local), t(wd["det.conv1.b"]).to(dtype))
        self.w["det2"] = (conv_w(wd["det.conv2.w"], 1, 128), t(wd["det.conv2.b"]).to(dtype))
        self.vm_w, self.vm_b = t(wd["vlad.memberships.w"]).to(dtype), t(wd["vlad.memberships.b"]).to(dtype)
        self.cl = t(wd["vlad.clusters"]).float()
        self.fc_w, self.fc_b = t(wd["fc.w"]).to(dtype), t(wd["fc.b"]).float()

    def _conv(self, x, name, k, s, groups=1):
        w, b = self.w[name]
        if k > 1:
            (t_, b_), (l_, r_) = _same_pad(x.shape[2], k, s), _same_pad(x.shape[3], k, s)
            x = F.pad(x, (l_, r_, t_, b_))
        return F.conv2d(x, w, b, stride=s, groups=groups)

    @torch.no_grad()
    def forward(self, img_u8):
        x = ((img_u8.to(self.dtype) - 128.0) / 128.0)[:, None].contiguous(memory_format=torch.channels_last)
        x = torch.clamp(self._conv(x, "conv1", 3, 2), 0.0, 6.0)
        local = None
        for b in self.blocks:
            p = f"l{b.layer}"
            inp = x
            if b.has_expand:
                x = torch.clamp(self._conv(x, p + ".expand", 1, 1), 0.0, 6.0)
            x = torch.clamp(self._conv(x, p + ".dw", 3, b.stride, groups=b.cexp), 0.0, 6.0)
            x = self._conv(x, p + ".project", 1, 1)
            if b.residual:
                x = x + inp
            if b.layer == LOCAL_ENDPOINT:
                local = x
        d = torch.clamp(self._conv(local, "desc1", 3, 1), 0.0, 6.0)
        d = self._conv(d, "desc2", 1, 1).permute(0, 2, 3, 1).float()
        d = d * torch.rsqrt(torch.clamp((d * d).sum(-1, keepdim=True), min=1e-12))
        l = torch.clamp(self._conv(local, "det1", 3, 1), 0.0, 6.0)
        l = self._conv(l, "det2", 1, 1).permute(0, 2, 3, 1).float()
        prob = torch.softmax(l, dim=-1)[..., :-1]
        B, Hc, Wc, _ = prob.shape
        g = DET_GRID
        prob = prob.reshape(B, Hc, Wc, g, g).permute(0, 1, 3, 2, 4).reshape(B, Hc * g, Wc * g)
        mp = lambda s: F.max_pool2d(s[:, None], kernel_size=9, stride=1, padding=4)[:, 0]
        zeros = torch.zeros_like(prob)
        mask = prob == mp(prob)
        supp = mp(mask.float()) > 0
        ss = torch.where(supp, zeros, prob)
        mask = mask | ((ss == mp(ss)) & ~supp)
        nms = torch.where(mask, prob, zeros)
        f = x.permute(0, 2, 3, 1)
        m = torch.softmax((torch.matmul(f, self.vm_w) + self.vm_b).float(), dim=-1)
        ff = f.float()
        v = m.sum(dim=(1, 2))[:, :, None] * self.cl[None] - torch.einsum("bhwc,bhwd->bcd", m, ff)
        v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
        v = v.reshape(B, -1)
        v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
        v = v * torch.rsqrt(torch.clamp((v * v).sum(1, keepdim=True), min=1e-12))
        y = torch.matmul(v.to(self.dtype), self.fc_w).float() + self.fc_b
        y = y * torch.rsqrt(torch.clamp((y * y).sum(1, keepdim=True), min=1e-12))
        return nms, d, y


def cudnn_fp16_forward_ms(wd, frames_u8: np.ndarray, dev, iters: int = 20):
    """ms per forward of the batch `frames_u8` [B,H,W]: (eager ms, CUDA-graph replay ms)."""
    torch.backends.cudnn.benchmark = True
    net = LibraryHFNet(wd, dev)
    H8, W8 = frames_u8.shape[1] // 8 * 8, frames_u8.shape[2] // 8 * 8
    x = torch.from_numpy(np.ascontiguousarray(frames_u8[:, :H8, :W8])).to(dev)
    for _ in range(3):
        net.forward(x)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net.forward(x)
    e1.record()
    torch.cuda.synchronize(dev)
    eager = e0.elapsed_time(e1) / iters
    graph_ms = None
    try:
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                net.forward(x)
        torch.cuda.current_stream(dev).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = net.forward(x)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        graph_ms = e0.elapsed_time(e1) / iters
    except Exception as ex:      # graph capture is best effort: the eager number stands
        graph_ms = None
    return eager, graph_ms
