"""CPU restatement of the reference's image preprocessing and prediction post-processing -- ORACLE only (test
infrastructure; nothing under spegnet_b200/ imports it).

* `process_image_array` follows CODImageProcessor.process_image (utils/image_processor.py:114-134) from the decoded RGB
  array on: float / 255 -> F.interpolate(size, bilinear, align_corners=False, antialias=True) -> (x - mean) / std.
  The file decode (PIL) stays with the caller.  Pinned against the reference class itself run on PNG fixtures
  (tests/golden/make_golden_preprocess.py -> preprocess.npz, tests/test_oracle_preprocess.py).
* `aa_resize_rows` is an explicit numpy restatement of ATen's anti-aliasing resampler for one axis
  (aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa), pinned against F.interpolate;
  it documents the index arithmetic the CUDA kernel (csrc/imageio.cu) reproduces.
* `resize_logits` follows engine/predictor.py:350-365 / engine/evaluator.py:539-544.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def process_image_array(rgb_u8: np.ndarray, target: int, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    img = torch.from_numpy(np.ascontiguousarray(rgb_u8)).float().permute(2, 0, 1) / 255.0
    img = F.interpolate(img.unsqueeze(0), size=(target, target), mode="bilinear", align_corners=False,
                        antialias=True).squeeze(0)
    return (img - torch.tensor(mean).view(-1, 1, 1)) / torch.tensor(std).view(-1, 1, 1)


def aa_windows(in_size: int, out_size: int):
    """(xmin, weights) per output index, float32 weights, with ATen's float/double promotions."""
    scale = np.float32(in_size) / np.float32(out_size)
    support = np.float32(1.0 * float(scale)) if scale >= 1.0 else np.float32(1.0)
    invscale = np.float32(1.0) / scale if scale >= 1.0 else np.float32(1.0)
    out = []
    for i in range(out_size):
        center = np.float32(float(scale) * (i + 0.5))
        xmin = max(int(float(center) - float(support) + 0.5), 0)
        xsize = min(int(float(center) + float(support) + 0.5), in_size) - xmin
        w = np.zeros(xsize, np.float32)
        for j in range(xsize):
            arg = np.float32((float(np.float32(j + xmin) - center) + 0.5) * float(invscale))
            a = abs(arg)
            w[j] = np.float32(1.0) - a if a < 1.0 else np.float32(0.0)
        total = np.float32(0.0)
        for v in w:
            total = np.float32(total + v)
        if total != 0:
            w = (w / total).astype(np.float32)
        out.append((xmin, w))
    return out


def aa_resize_rows(x: np.ndarray, out_size: int) -> np.ndarray:
    """Resample the LAST axis of a float32 array to out_size (one pass of the separable resampler)."""
    if x.shape[-1] == out_size:
        return x.copy()
    res = np.zeros(x.shape[:-1] + (out_size,), np.float32)
    for i, (xmin, w) in enumerate(aa_windows(x.shape[-1], out_size)):
        acc = x[..., xmin] * w[0]
        for j in range(1, len(w)):
            acc = (acc + x[..., xmin + j] * w[j]).astype(np.float32)
        res[..., i] = acc
    return res


def resize_logits(logits: torch.Tensor, size, sigmoid: bool = True) -> torch.Tensor:
    out = F.interpolate(logits, size=size, mode="bilinear", align_corners=False)
    return out.sigmoid() if sigmoid else out
