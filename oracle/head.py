"""fp32 CPU restatement of SPEGNet's CFI / EFE / PED head -- ORACLE, test infrastructure only.

Written functionally over a flat state dict with the reference's parameter names, not as a copy of the
reference modules; each function cites the reference lines whose arithmetic it restates.  It is pinned
against the reference modules imported verbatim (tests/golden/make_golden.py -> tests/golden/head_*.npz,
checked by tests/test_oracle_head.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # nn.BatchNorm2d default; the reference never overrides it


def _bn(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"],
                        sd[pre + "bias"], training=False, eps=BN_EPS)


def _up(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def fusion(sd: SD, pre: str, f2: torch.Tensor, f3: torch.Tensor, f4: torch.Tensor) -> torch.Tensor:
    """AdaptiveAttentionFusion (models/feature_integration.py:205-246) + SE (:128-151)."""
    size = f2.shape[-2:]
    cat = torch.cat([f2, _up(f3, size), _up(f4, size)], dim=1)
    x = F.relu(_bn(sd, pre + "bn.", F.conv2d(cat, sd[pre + "conv1x1.weight"])))
    squeeze = x.mean(dim=(2, 3))
    gate = torch.sigmoid(F.linear(F.relu(F.linear(squeeze, sd[pre + "se_block.fc.0.weight"])),
                                  sd[pre + "se_block.fc.2.weight"]))
    return x * gate[:, :, None, None]


def easpp(sd: SD, pre: str, x: torch.Tensor, rates=(1, 6, 12, 18)) -> torch.Tensor:
    """EfficientASPP (models/feature_integration.py:369-417)."""
    size = x.shape[-2:]
    r = F.relu(_bn(sd, pre + "reduce.1.", F.conv2d(x, sd[pre + "reduce.0.weight"])))
    c = r.shape[1]
    branches: List[torch.Tensor] = []
    for i, d in enumerate(rates):
        y = F.conv2d(r, sd[f"{pre}branches.{i}.0.weight"], padding=d, dilation=d, groups=c)
        branches.append(F.relu(_bn(sd, f"{pre}branches.{i}.1.", y)))
    g = r.mean(dim=(2, 3), keepdim=True)
    g = F.relu(_bn(sd, pre + "global_branch.2.", F.conv2d(g, sd[pre + "global_branch.1.weight"])))
    branches.append(g.expand(-1, -1, *size))  # bilinear 1x1 -> HxW is a constant broadcast (:402-407)
    cat = torch.cat(branches, dim=1)
    y = F.relu(_bn(sd, pre + "fusion.1.", F.conv2d(cat, sd[pre + "fusion.0.weight"], groups=c)))
    return F.relu(_bn(sd, pre + "expand.1.", F.conv2d(y, sd[pre + "expand.0.weight"])))


def edge_detector(sd: SD, pre: str, ctx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """EdgeDetectionModule (models/object_detection.py:132-157): returns (edge_map, edge_features)."""
    feat = F.relu(_bn(sd, pre + "bn1.", F.conv2d(ctx, sd[pre + "conv1.weight"], padding=1)))
    return F.conv2d(feat, sd[pre + "edge_conv.weight"], sd[pre + "edge_conv.bias"]), feat


def decoder(sd: SD, pre: str, ctx: torch.Tensor, edges: List[torch.Tensor | None]) -> List[torch.Tensor]:
    """BoundaryAwareDecoder + DecoderBlock (models/object_detection.py:201-238,309-342)."""
    x, preds = ctx, []
    for i, edge in enumerate(edges):
        b = f"{pre}decoder_blocks.{i}."
        x = _up(x, (2 * x.shape[-2], 2 * x.shape[-1]))
        if edge is not None:
            x = torch.cat([x, _up(edge, x.shape[-2:])], dim=1)
        x = F.relu(_bn(sd, b + "bn1.", F.conv2d(x, sd[b + "conv1.weight"], sd[b + "conv1.bias"], padding=1)))
        x = F.relu(_bn(sd, b + "bn2.", F.conv2d(x, sd[b + "conv2.weight"], sd[b + "conv2.bias"], padding=1)))
        preds.append(F.conv2d(x, sd[f"{pre}pred_heads.{i}.weight"], sd[f"{pre}pred_heads.{i}.bias"]))
    return preds


def head_forward(sd: SD, feats: List[torch.Tensor]) -> Dict[str, object]:
    """Everything after the encoder in SPEGNet.forward (models/spegnet.py:168-206).
    `feats` = the 4 encoder maps; stage 1 is unused, exactly as in the reference (:169-171)."""
    fused = fusion(sd, "fusion.", feats[1], feats[2], feats[3])
    ctx = easpp(sd, "context.", fused)
    edge_map, edge_feat = edge_detector(sd, "edge_detector.", ctx)
    preds = decoder(sd, "decoder.", ctx, [edge_feat, edge_feat, None])
    return {"predictions": preds, "edge": edge_map,
            "features": {"context": ctx, "fused": fused, "edge_features": edge_feat}}
