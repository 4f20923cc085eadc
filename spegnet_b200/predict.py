"""Single-image prediction path of the reference's ``engine/predictor.py`` with every stage on the GPU.

``Predictor.predict_single`` (engine/predictor.py:311-368) = ``CODImageProcessor.process_image`` (decode, /255,
antialiased bilinear resize to target_size, ImageNet normalisation: utils/image_processor.py:114-134) ->
``model(x)`` -> ``F.interpolate`` of the finest logits and the edge logits to ``output_size`` -> ``.sigmoid()``;
``save_binary_visualization`` then writes ``uint8(pred * 255)`` (utils/visualization.py:92-115).

Here the decoded uint8 RGB array is uploaded once and `spg_preprocess_rgb_u8` -> forward -> `spg_resize_bilinear_f32`
(+ sigmoid) run back to back on the current stream; the file decode / encode stays with the caller (PIL / cv2 in the
reference).  There is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


@torch.no_grad()
def preprocess(rgb_u8: torch.Tensor, target_size: int) -> torch.Tensor:
    """uint8 [H,W,3] on the device -> fp32 [1,3,target,target] model input."""
    if not rgb_u8.is_cuda:
        raise RuntimeError("spegnet_b200.predict runs on a CUDA (B200) device only; there is no CPU fallback")
    return ops.preprocess_rgb(rgb_u8.contiguous(), target_size)[None]


@torch.no_grad()
def predict_single(model, rgb_u8: torch.Tensor, target_size: int = 512,
                   output_size: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(segmentation probability [h,w], edge probability [h,w]) as fp32 device tensors, `output_size` = (h, w) or the
    model's native sizes when None (engine/predictor.py:344-365)."""
    out = model(preprocess(rgb_u8, target_size))
    seg, edge = out["predictions"][-1], out["edge"]
    seg_size = tuple(output_size) if output_size else tuple(seg.shape[-2:])
    edge_size = tuple(output_size) if output_size else tuple(edge.shape[-2:])
    return (ops.resize_bilinear(seg[0], seg_size, sigmoid=True)[0], ops.resize_bilinear(edge[0], edge_size, sigmoid=True)[0])


def binary_mask_u8(prob: torch.Tensor) -> torch.Tensor:
    """``(prediction * 255).astype(np.uint8)`` of save_binary_visualization (utils/visualization.py:109), on the device."""
    return (prob * 255).to(torch.uint8)
