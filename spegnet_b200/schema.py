"""Geometry and parameter schema of SPEGNet (Hiera-L trunk + CFI / EFE / PED head).

Product-side statement of what `models/spegnet.py:90-135` builds: the state-dict key names and shapes
(so reference checkpoints load unchanged) and the per-block geometry of the sam2 Hiera-L trunk
(`configs/default.yaml:4` -> sam2.1_hiera_l.yaml; HF:modeling_sam2.py:452-470).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

# models/feature_encoding.py:141-148 -- variants the reference names; only 'large' is wired to a trunk here.
VARIANT_CHANNELS = {
    "tiny": [96, 192, 384, 768], "small": [96, 192, 384, 768], "base": [96, 192, 384, 768],
    "base_plus": [112, 224, 448, 896], "large": [144, 288, 576, 1152], "huge": [256, 512, 1024, 2048],
}

HEAD_DIM = 72
LN_EPS = 1e-6
BN_EPS = 1e-5
ASPP_DILATIONS = (1, 6, 12, 18)  # models/spegnet.py:115-120


@dataclass(frozen=True)
class TrunkSpec:
    embed_dim: int = 144
    num_heads: int = 2
    stages: Tuple[int, ...] = (2, 6, 36, 4)
    global_att_blocks: Tuple[int, ...] = (23, 33, 43)
    window_spec: Tuple[int, ...] = (8, 4, 16, 8)
    pos_embed_bkg: Tuple[int, int] = (7, 7)
    q_pool: int = 3
    mlp_ratio: int = 4

    @property
    def dims(self) -> List[int]:
        return [self.embed_dim << s for s in range(len(self.stages))]

    @property
    def heads(self) -> List[int]:
        return [self.num_heads << s for s in range(len(self.stages))]

    @property
    def stage_ends(self) -> List[int]:
        out, acc = [], 0
        for n in self.stages:
            acc += n
            out.append(acc - 1)
        return out


@dataclass(frozen=True)
class Block:
    index: int
    stage: int
    dim_in: int
    dim_out: int
    heads: int
    window: int   # 0 = global
    q_pool: bool  # queries (and the shortcut) are 2x2 max-pooled: resolution halves in this block


def trunk_blocks(spec: TrunkSpec = TrunkSpec()) -> List[Block]:
    blocks: List[Block] = []
    i = 0
    for s, depth in enumerate(spec.stages):
        for j in range(depth):
            first = j == 0 and s > 0
            window = spec.window_spec[s - 1] if first else spec.window_spec[s]  # window lags one block
            if i in spec.global_att_blocks:
                window = 0
            blocks.append(Block(i, s, spec.dims[s - 1] if first else spec.dims[s], spec.dims[s], spec.heads[s],
                                window, first and s <= spec.q_pool))
            i += 1
    return blocks


Shape = Tuple[int, ...]


def trunk_param_shapes(spec: TrunkSpec = TrunkSpec()) -> Dict[str, Shape]:
    d0 = spec.embed_dim
    out: Dict[str, Shape] = {
        "patch_embed.proj.weight": (d0, 3, 7, 7), "patch_embed.proj.bias": (d0,),
        "pos_embed": (1, d0) + tuple(spec.pos_embed_bkg),
        "pos_embed_window": (1, d0, spec.window_spec[0], spec.window_spec[0]),
    }
    for b in trunk_blocks(spec):
        p = f"blocks.{b.index}."
        hid = b.dim_out * spec.mlp_ratio
        out[p + "norm1.weight"] = (b.dim_in,)
        out[p + "norm1.bias"] = (b.dim_in,)
        out[p + "attn.qkv.weight"] = (3 * b.dim_out, b.dim_in)
        out[p + "attn.qkv.bias"] = (3 * b.dim_out,)
        out[p + "attn.proj.weight"] = (b.dim_out, b.dim_out)
        out[p + "attn.proj.bias"] = (b.dim_out,)
        out[p + "norm2.weight"] = (b.dim_out,)
        out[p + "norm2.bias"] = (b.dim_out,)
        out[p + "mlp.layers.0.weight"] = (hid, b.dim_out)
        out[p + "mlp.layers.0.bias"] = (hid,)
        out[p + "mlp.layers.1.weight"] = (b.dim_out, hid)
        out[p + "mlp.layers.1.bias"] = (b.dim_out,)
        if b.dim_in != b.dim_out:
            out[p + "proj.weight"] = (b.dim_out, b.dim_in)
            out[p + "proj.bias"] = (b.dim_out,)
    return out


def _bn(out: Dict[str, Tuple[Shape, str]], prefix: str, c: int) -> None:
    out[prefix + "weight"] = ((c,), "param")
    out[prefix + "bias"] = ((c,), "param")
    out[prefix + "running_mean"] = ((c,), "buffer")
    out[prefix + "running_var"] = ((c,), "buffer")
    out[prefix + "num_batches_tracked"] = ((), "buffer_long")


def head_entries(enc: Tuple[int, int, int] = (288, 576, 1152)) -> Dict[str, Tuple[Shape, str]]:
    """name -> (shape, 'param' | 'buffer' | 'buffer_long') for fusion / context / edge_detector / decoder
    (models/feature_integration.py:196-203,309-367; models/object_detection.py:108-130,183-198,295-307)."""
    out: Dict[str, Tuple[Shape, str]] = {}
    P = "param"
    out["fusion.conv1x1.weight"] = ((512, sum(enc), 1, 1), P)
    _bn(out, "fusion.bn.", 512)
    out["fusion.se_block.fc.0.weight"] = ((32, 512), P)
    out["fusion.se_block.fc.2.weight"] = ((512, 32), P)
    out["context.reduce.0.weight"] = ((128, 512, 1, 1), P)
    _bn(out, "context.reduce.1.", 128)
    for i in range(4):
        out[f"context.branches.{i}.0.weight"] = ((128, 1, 3, 3), P)
        _bn(out, f"context.branches.{i}.1.", 128)
    out["context.global_branch.1.weight"] = ((128, 128, 1, 1), P)
    _bn(out, "context.global_branch.2.", 128)
    out["context.fusion.0.weight"] = ((128, 5, 1, 1), P)
    _bn(out, "context.fusion.1.", 128)
    out["context.expand.0.weight"] = ((256, 128, 1, 1), P)
    _bn(out, "context.expand.1.", 256)
    out["edge_detector.conv1.weight"] = ((64, 256, 3, 3), P)
    _bn(out, "edge_detector.bn1.", 64)
    out["edge_detector.edge_conv.weight"] = ((1, 64, 1, 1), P)
    out["edge_detector.edge_conv.bias"] = ((1,), P)
    cin, cout = (320, 320, 128), (256, 128, 64)
    for i in range(3):
        p = f"decoder.decoder_blocks.{i}."
        out[p + "conv1.weight"] = ((cout[i], cin[i], 3, 3), P)
        out[p + "conv1.bias"] = ((cout[i],), P)
        _bn(out, p + "bn1.", cout[i])
        out[p + "conv2.weight"] = ((cout[i], cout[i], 3, 3), P)
        out[p + "conv2.bias"] = ((cout[i],), P)
        _bn(out, p + "bn2.", cout[i])
    for i in range(3):
        out[f"decoder.pred_heads.{i}.weight"] = ((1, cout[i], 1, 1), P)
        out[f"decoder.pred_heads.{i}.bias"] = ((1,), P)
    return out
