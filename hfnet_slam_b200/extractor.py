"""Host-side mirror of the reference's extractor interface over the C-ABI:

* ``BaseModel`` / ``HFNetB200Model``   include/Extractors/BaseModel.h:38-54 (Detect overloads, IsValid, Type)
* ``init_all_models`` / ``get_model_vec`` src/Extractors/BaseModel.cc:24-113 (process-global registry, one model per level)
* ``HFextractor``                       include/Extractors/HFextractor.h, src/Extractors/HFextractor.cc:21-157

Unlike the reference (one TensorRT engine and one synchronous H2D/D2H pair per pyramid level, HFextractor.cc:255-284)
all levels of a frame run in one context as one CUDA graph with one H2D and one D2H burst.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .lib import Context, HfbError

K_HFNET_B200_MODEL = 3            # extends enum ModelType {kHFNetTFModel, kHFNetRTModel, kHFNetVINOModel}
K_IMAGE_TO_LOCAL_AND_GLOBAL, K_IMAGE_TO_LOCAL = 0, 1


def features_per_level(nfeatures: int, nlevels: int, scale_factor: float) -> List[int]:
    """src/Extractors/HFextractor.cc:108-119 (float arithmetic, cvRound)."""
    factor = np.float32(1.0) / np.float32(scale_factor)
    n = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(float(factor) ** nlevels))
    out, total = [], 0
    for _ in range(nlevels - 1):
        v = int(np.rint(np.float32(n)))
        out.append(v)
        total += v
        n = np.float32(n * factor)
    out.append(max(nfeatures - total, 0))
    return out


class KeyPoint:
    """cv::KeyPoint fields the reference fills (HFNetRTModel.cc:150-166)."""
    __slots__ = ("x", "y", "response", "octave", "angle")

    def __init__(self, x, y, response, octave):
        self.x, self.y, self.response, self.octave, self.angle = float(x), float(y), float(response), int(octave), 0.0


class HFNetB200Model:
    """One pyramid level's facade over a shared context (the reference keeps one engine per level)."""

    def __init__(self, ctx: Context, level: int, mode: int):
        self.ctx, self.level, self.mode = ctx, level, mode

    def is_valid(self) -> bool:
        return bool(self.ctx.handle)

    def type(self) -> int:
        return K_HFNET_B200_MODEL

    def detect(self, image: np.ndarray, n_keypoints: int, threshold: float, want_global: Optional[bool] = None):
        """bool Detect(image, vKeyPoints, localDescriptors[, globalDescriptors], nKeypointsNum, threshold).
        Returns (ok, keypoints dict, local descriptors [N,256], global descriptor [4096,1] or None).  Like the
        reference it returns ok=False on a mode mismatch instead of raising (HFNetRTModel.cc:87-91)."""
        if want_global is None:
            want_global = self.mode == K_IMAGE_TO_LOCAL_AND_GLOBAL
        if want_global and self.mode != K_IMAGE_TO_LOCAL_AND_GLOBAL:
            return False, None, None, None
        if self.level != 0 or self.ctx.n_levels != 1:
            raise HfbError(1, "per-level Detect needs a single-level context; use HFextractor for pyramids")
        out = self.ctx.extract(image, [n_keypoints], threshold)
        g = out["global_descriptor"].reshape(4096, 1) if want_global and out["global_descriptor"] is not None else None
        return True, out, out["descriptors"], g


_models: List[HFNetB200Model] = []
_ctx: Optional[Context] = None


def init_all_models(weights_blob: bytes, image_size: Tuple[int, int], n_levels: int, scale_factor: float,
                    max_keypoints: int = 4096, max_batch: int = 1, device: int = 0) -> List[HFNetB200Model]:
    """InitAllModels(strModelPath, modelType, ImSize, nLevels, scaleFactor) (src/Extractors/BaseModel.cc:24-93):
    level 0 is kImageToLocalAndGlobal, the others kImageToLocal.  image_size = (width, height) like cv::Size."""
    global _models, _ctx
    if _ctx is not None:
        _ctx.close()
    w, h = image_size
    _ctx = Context(height=h, width=w, n_levels=n_levels, scale_factor=scale_factor, max_keypoints=max_keypoints,
                   max_batch=max_batch, with_global=True, device=device)
    _ctx.load_weights(weights_blob)
    _models = [HFNetB200Model(_ctx, l, K_IMAGE_TO_LOCAL_AND_GLOBAL if l == 0 else K_IMAGE_TO_LOCAL)
               for l in range(n_levels)]
    return _models


def get_model_vec() -> List[HFNetB200Model]:
    return _models


def get_global_model():
    """nullptr for the TensorRT back-end too (src/Extractors/BaseModel.cc:78-81): level 0 already yields the global
    descriptor."""
    return None


class HFextractor:
    """HFextractor(nfeatures, threshold, scaleFactor, nlevels, vpModels) (src/Extractors/HFextractor.cc:21-79)."""

    def __init__(self, nfeatures: int, threshold: float, scale_factor: float, nlevels: int,
                 models: Sequence[HFNetB200Model]):
        if len(models) != nlevels:
            raise HfbError(1, "one model per level expected")
        self.nfeatures, self.threshold, self.scale_factor, self.nlevels = nfeatures, threshold, scale_factor, nlevels
        self.ctx = models[0].ctx
        if self.ctx.n_levels != nlevels:
            raise HfbError(1, "context was created for a different number of levels")
        sf = [np.float32(1.0)]
        for _ in range(1, nlevels):
            sf.append(np.float32(sf[-1] * np.float32(scale_factor)))
        self.mvScaleFactor = sf
        self.mvInvScaleFactor = [np.float32(1.0) / s for s in sf]
        self.mvLevelSigma2 = [s * s for s in sf]
        self.mvInvLevelSigma2 = [np.float32(1.0) / s for s in self.mvLevelSigma2]
        self.mnFeaturesPerLevel = features_per_level(nfeatures, nlevels, scale_factor) if nlevels > 1 else [nfeatures]

    def __call__(self, image: np.ndarray):
        """int operator()(image, vKeyPoints, localDescriptors, globalDescriptors) (HFextractor.cc:142-157): returns
        (n, keypoints dict, local descriptors [N,256], global descriptor [4096,1]); n = -1 on a bad image
        (empty or not CV_8UC1, :145)."""
        img = np.asarray(image)
        if img.size == 0 or img.dtype != np.uint8 or img.ndim != 2:
            return -1, None, None, None
        out = self.ctx.extract(img, self.mnFeaturesPerLevel, self.threshold)
        return len(out["x"]), out, out["descriptors"], out["global_descriptor"].reshape(4096, 1)

    def extract_batch(self, images: Sequence[np.ndarray]):
        return self.ctx.extract_batch(images, self.mnFeaturesPerLevel, self.threshold)
