"""Whole-model fp32 CPU oracle: Hiera trunk + head, driven by a reference-schema state dict.
ORACLE -- test infrastructure only (see oracle/__init__.py).

Restates SPEGNet.forward (models/spegnet.py:137-206) and the input checks of
HieraSAM2FeatureEncoder.forward (models/feature_encoding.py:230-233).
"""
from __future__ import annotations

from typing import Dict

import torch

from .head import head_forward
from .hiera import HieraConfig, hiera_forward

TRUNK_PREFIX = "encoder.encoder."  # models/spegnet.py:94 + models/feature_encoding.py:159


@torch.inference_mode()
def spegnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: HieraConfig = HieraConfig(),
                    taps: Dict[str, torch.Tensor] | None = None) -> Dict[str, object]:
    if x.dim() != 4:
        raise ValueError(f"Expected 4D input (B,C,H,W), got {x.dim()}D")
    if any(s % 32 != 0 for s in x.shape[-2:]):
        raise ValueError("Input spatial dims must be divisible by 32")
    feats = hiera_forward(sd, x.float(), cfg, prefix=TRUNK_PREFIX, taps=taps)
    out = head_forward(sd, feats)
    if taps is not None:
        for i, f in enumerate(feats):
            taps[f"stage{i + 1}"] = f
    return out
