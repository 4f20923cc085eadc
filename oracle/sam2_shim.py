"""`sys.modules` shim for the absent third-party `sam2` package -- ORACLE, test infrastructure only.

The reference does `from sam2.build_sam import build_sam2` (models/feature_encoding.py:107) and keeps
`build_sam2(...).image_encoder.trunk` (:156-159).  With this shim installed, `models/spegnet.py` of
the reference imports and constructs VERBATIM around the restated trunk (oracle/hiera.py); it is used in
the dev container only (tests/golden/make_golden.py) -- /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import sys
import types

import torch.nn as nn

from .hiera import HieraConfig, HieraTrunk


class _ImageEncoder(nn.Module):
    def __init__(self, cfg: HieraConfig):
        super().__init__()
        self.trunk = HieraTrunk(cfg)


class _Sam2Model(nn.Module):
    """Has `image_encoder` plus one throw-away child so _clear_unused_components has something to delete
    (models/feature_encoding.py:186-196)."""

    def __init__(self, cfg: HieraConfig):
        super().__init__()
        self.image_encoder = _ImageEncoder(cfg)
        self.memory_attention = nn.Identity()


def install(cfg: HieraConfig = HieraConfig()) -> None:
    def build_sam2(config_file=None, ckpt_path=None, device="cuda", mode="eval",
                   hydra_overrides_extra=(), apply_postprocessing=True, **_):
        return _Sam2Model(cfg)

    pkg = types.ModuleType("sam2")
    sub = types.ModuleType("sam2.build_sam")
    sub.build_sam2 = build_sam2
    pkg.build_sam = sub
    sys.modules["sam2"] = pkg
    sys.modules["sam2.build_sam"] = sub
