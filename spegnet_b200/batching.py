"""Batch assembly of the reference's data path, for ragged inputs.

``utils/data_loader.py::collate_fn`` (:177-212) stacks the (already resized) images and keeps masks / edges as LISTS,
because they stay at their original, per-image resolution (``CODImageProcessor.process_mask`` only thresholds at 127.5,
utils/image_processor.py:140-172); the evaluator then resizes every prediction to its own mask
(engine/evaluator.py:539-544).

* ``collate_fn`` is that function with the same keys, ordering and error, for samples produced on the host.
* ``collate_decoded`` starts one step earlier, from DECODED uint8 arrays of different sizes already on the device: every
  image goes through ``spg_preprocess_rgb_u8`` (the /255 + antialiased resize + ImageNet normalisation of
  ``process_image``) straight into its slot of one [B,3,S,S] batch tensor -- no per-image host tensor, no stack copy --
  and the masks are thresholded on the device and stay a ragged list.  File decode stays with the caller (PIL / cv2 in
  the reference).  ``spegnet_b200.evaluate.score_batch`` consumes the result as is.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops


def collate_fn(batch: List[Dict]) -> Dict:
    """utils/data_loader.py:177-212: {'images': [B,C,H,W], 'masks': list of [1,h_i,w_i]} plus 'edges' (training
    samples) or 'names' (test samples); ``ValueError`` on an empty batch."""
    if not batch:
        raise ValueError("Empty batch received")
    out = {"images": torch.stack([item["image"] for item in batch]), "masks": [item["mask"] for item in batch]}
    if "edge" in batch[0]:
        out["edges"] = [item["edge"] for item in batch]
    else:
        out["names"] = [item["name"] for item in batch]
    return out


@torch.no_grad()
def collate_decoded(images_u8: Sequence[torch.Tensor], masks_u8: Optional[Sequence[torch.Tensor]] = None,
                    names: Optional[Sequence[str]] = None, target_size: int = 512) -> Dict:
    """images_u8: uint8 [h_i, w_i, 3] device tensors of any sizes; masks_u8: uint8 [H_i, W_i] grey-level masks (device).
    Returns the same dictionary layout as `collate_fn` with 'images' fp32 [B,3,S,S] on the device and 'masks' a list of
    fp32 {0,1} [1,H_i,W_i] tensors (``(mask > 127.5).float()``, utils/image_processor.py:161-166)."""
    if len(images_u8) == 0:
        raise ValueError("Empty batch received")
    dev = images_u8[0].device
    if dev.type != "cuda":
        raise RuntimeError("spegnet_b200.batching.collate_decoded runs on a CUDA (B200) device only; there is no CPU fallback")
    images = torch.empty(len(images_u8), 3, target_size, target_size, dtype=torch.float32, device=dev)
    for i, img in enumerate(images_u8):
        ops.preprocess_rgb(img.contiguous(), target_size, out=images[i])
    out: Dict = {"images": images}
    if masks_u8 is not None:
        if len(masks_u8) != len(images_u8):
            raise ValueError(f"{len(images_u8)} images but {len(masks_u8)} masks")
        out["masks"] = [(m > 127).to(torch.float32)[None] for m in masks_u8]  # uint8 > 127.5  <=>  > 127
    if names is not None:
        out["names"] = list(names)
    return out
