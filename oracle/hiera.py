"""fp32 CPU restatement of the SAM2 Hiera trunk -- ORACLE, test infrastructure only.

The trunk is third-party code (`facebookresearch/sam2`, installed from an unpinned git HEAD by the
reference's setup/environment.yml:25) and is NOT in /root/reference: the reference only calls it
(models/feature_encoding.py:107,156,159,236).  Parity with upstream sam2 is therefore **unpinned**;
this restatement follows the published Hiera/"hieradet" algorithm and is pinned in
tests/test_oracle_trunk.py against the independent HF port that ships in this image
(transformers/models/sam2/modeling_sam2.py:119-149 patch embed, :307-345 attention, :378-438 window
partition, :441-531 block, :588-654 model).

Everything is driven by a flat state dict that uses the upstream parameter names
(``patch_embed.proj.*``, ``pos_embed``, ``pos_embed_window``, ``blocks.{i}.norm1|attn.qkv|attn.proj|
norm2|mlp.layers.{0,1}|proj``), so a real SPEGNet checkpoint (``encoder.encoder.<name>``) maps 1:1.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class HieraConfig:
    """Hyper-parameters of sam2.1_hiera_l.yaml (configs/default.yaml:4 points at it)."""

    embed_dim: int = 144
    num_heads: int = 2
    stages: Tuple[int, ...] = (2, 6, 36, 4)
    global_att_blocks: Tuple[int, ...] = (23, 33, 43)
    window_spec: Tuple[int, ...] = (8, 4, 16, 8)
    window_pos_embed_bkg_spatial_size: Tuple[int, int] = (7, 7)
    q_pool: int = 3
    q_stride: int = 2
    dim_mul: float = 2.0
    head_mul: float = 2.0
    mlp_ratio: float = 4.0
    ln_eps: float = 1e-6

    @property
    def depth(self) -> int:
        return sum(self.stages)

    @property
    def stage_ends(self) -> List[int]:
        ends, acc = [], 0
        for n in self.stages:
            acc += n
            ends.append(acc - 1)
        return ends

    @property
    def stage_dims(self) -> List[int]:
        return [int(self.embed_dim * self.dim_mul ** s) for s in range(len(self.stages))]

    @property
    def stage_heads(self) -> List[int]:
        return [int(self.num_heads * self.head_mul ** s) for s in range(len(self.stages))]


@dataclass(frozen=True)
class BlockSpec:
    index: int
    stage: int
    dim_in: int
    dim_out: int
    heads: int
    window: int  # 0 = global attention
    q_stride: int  # 0 = no query pooling


def block_specs(cfg: HieraConfig) -> List[BlockSpec]:
    """Per-block geometry.  The window size lags the stage change by one block (the first block of a
    stage still partitions with the previous stage's window) and the first block of stages
    1..q_pool pools its queries 2x2 (HF:modeling_sam2.py:452-470)."""
    dims, heads = cfg.stage_dims, cfg.stage_heads
    specs: List[BlockSpec] = []
    idx = 0
    for s, depth in enumerate(cfg.stages):
        for j in range(depth):
            first = j == 0 and s > 0
            window = cfg.window_spec[s - 1] if first else cfg.window_spec[s]
            if idx in cfg.global_att_blocks:
                window = 0
            specs.append(
                BlockSpec(
                    index=idx,
                    stage=s,
                    dim_in=dims[s - 1] if first else dims[s],
                    dim_out=dims[s],
                    heads=heads[s],
                    window=window,
                    q_stride=cfg.q_stride if (first and s <= cfg.q_pool) else 0,
                )
            )
            idx += 1
    return specs


ATTENTION_IMPL = "einsum"  # "sdpa": F.scaled_dot_product_attention (library baseline on the GPU)


def _pool2x2_nhwc(t: torch.Tensor, stride: int) -> torch.Tensor:
    return F.max_pool2d(t.permute(0, 3, 1, 2), kernel_size=stride, stride=stride).permute(0, 2, 3, 1)


def _to_windows(t: torch.Tensor, ws: int):
    """[B,h,w,C] -> ([B*nW, ws, ws, C], (hp, wp)).  A grid that does not tile into ws x ws windows is padded with ZERO
    tokens at the bottom / right first (HF:modeling_sam2.py:395-399): the padding happens after norm1, so the padded
    tokens reach the qkv projection as zeros, come out as its bias, and take part in the window's softmax as keys."""
    b, h, w, c = t.shape
    ph, pw = (ws - h % ws) % ws, (ws - w % ws) % ws
    if ph or pw:
        t = F.pad(t, (0, 0, 0, pw, 0, ph))
    hp, wp = h + ph, w + pw
    t = t.reshape(b, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(b * (hp // ws) * (wp // ws), ws, ws, c), (hp, wp)


def _from_windows(t: torch.Tensor, ws: int, b: int, hp: int, wp: int, h: int, w: int) -> torch.Tensor:
    """Inverse of `_to_windows` on the padded grid, then crop to the h x w valid tokens (HF:...:417-438)."""
    c = t.shape[-1]
    t = t.reshape(b, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(b, hp, wp, c)[:, :h, :w]


def pos_embed_map(pos_embed: torch.Tensor, pos_embed_window: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """bicubic(background 7x7 -> h x w) + tiled window embedding, NHWC [1,h,w,C] (HF:...:623-629)."""
    bg = F.interpolate(pos_embed, size=(h, w), mode="bicubic")
    wh, ww = pos_embed_window.shape[-2:]
    tiled = pos_embed_window.repeat(1, 1, h // wh, w // ww)
    return (bg + tiled).permute(0, 2, 3, 1).contiguous()


def block_forward(sd: Dict[str, torch.Tensor], pre: str, x: torch.Tensor, spec: BlockSpec, eps: float) -> torch.Tensor:
    b, h, w, _ = x.shape
    hd = spec.dim_out // spec.heads
    y = F.layer_norm(x, (spec.dim_in,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    skip = x
    if spec.dim_in != spec.dim_out:
        # the shortcut is projected from the *normalised* tokens, then pooled
        skip = _pool2x2_nhwc(F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"]), spec.q_stride)

    ws = spec.window if spec.window > 0 else h
    if spec.window == 0 and h != w:
        raise ValueError("global attention restated for square token grids only")
    tokens, (hp, wp) = _to_windows(y, ws)  # [nW, ws, ws, C]
    nw = tokens.shape[0]
    qkv = F.linear(tokens.reshape(nw, ws * ws, spec.dim_in), sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(nw, ws * ws, 3, spec.heads, hd)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ws_out = ws
    if spec.q_stride:
        q = _pool2x2_nhwc(q.reshape(nw, ws, ws, spec.dim_out), spec.q_stride)
        ws_out = ws // spec.q_stride
        q = q.reshape(nw, ws_out * ws_out, spec.heads, hd)
    if ATTENTION_IMPL == "sdpa":
        # upstream sam2 calls F.scaled_dot_product_attention; only the GPU library-baseline leg of bench.py selects it
        # (the fused kernel is what the reference would run on a GPU); the CPU oracle keeps the explicit softmax
        ctx = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
        ctx = ctx.reshape(nw, ws_out, ws_out, spec.dim_out)
    else:
        scores = torch.einsum("wqhd,wkhd->whqk", q, k) * (1.0 / math.sqrt(hd))
        probs = scores.softmax(dim=-1)
        ctx = torch.einsum("whqk,wkhd->wqhd", probs, v).reshape(nw, ws_out, ws_out, spec.dim_out)
    ctx = F.linear(ctx, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    ho, wo = (h // spec.q_stride, w // spec.q_stride) if spec.q_stride else (h, w)
    hpo, wpo = (hp // spec.q_stride, wp // spec.q_stride) if spec.q_stride else (hp, wp)  # HF:...:514-521
    x = skip + _from_windows(ctx, ws_out, b, hpo, wpo, ho, wo)

    z = F.layer_norm(x, (spec.dim_out,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    z = F.linear(z, sd[pre + "mlp.layers.0.weight"], sd[pre + "mlp.layers.0.bias"])
    z = F.gelu(z)  # exact erf GELU
    z = F.linear(z, sd[pre + "mlp.layers.1.weight"], sd[pre + "mlp.layers.1.bias"])
    return x + z


def hiera_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: HieraConfig = HieraConfig(),
                  prefix: str = "", taps: Dict[str, torch.Tensor] | None = None) -> List[torch.Tensor]:
    """x: [B,3,S,S] fp32 -> 4 NCHW feature maps (strides 4/8/16/32).  `taps`, if given, receives the
    NHWC token tensor after the patch embed and after every block (for layer-wise GPU parity)."""
    if x.dim() != 4:
        raise ValueError(f"Expected 4D input (B,C,H,W), got {x.dim()}D")
    t = F.conv2d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=4, padding=3)
    t = t.permute(0, 2, 3, 1)
    t = t + pos_embed_map(sd[prefix + "pos_embed"], sd[prefix + "pos_embed_window"], t.shape[1], t.shape[2])
    if taps is not None:
        taps["embed"] = t
    outs: List[torch.Tensor] = []
    ends = cfg.stage_ends
    for spec in block_specs(cfg):
        t = block_forward(sd, f"{prefix}blocks.{spec.index}.", t, spec, cfg.ln_eps)
        if taps is not None:
            taps[f"block{spec.index}"] = t
        if spec.index in ends:
            outs.append(t.permute(0, 3, 1, 2))
    return outs


class HieraTrunk(nn.Module):
    """nn.Module carrier of the trunk parameters under the upstream names (so that the reference
    SPEGNet's state dict reads ``encoder.encoder.blocks.0.attn.qkv.weight`` etc.)."""

    def __init__(self, cfg: HieraConfig = HieraConfig()):
        super().__init__()
        self.cfg = cfg
        d0 = cfg.embed_dim
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, d0, kernel_size=7, stride=4, padding=3)
        self.pos_embed = nn.Parameter(torch.zeros(1, d0, *cfg.window_pos_embed_bkg_spatial_size))
        self.pos_embed_window = nn.Parameter(torch.zeros(1, d0, cfg.window_spec[0], cfg.window_spec[0]))
        self.blocks = nn.ModuleList()
        for spec in block_specs(cfg):
            blk = nn.Module()
            blk.norm1 = nn.LayerNorm(spec.dim_in, eps=cfg.ln_eps)
            blk.attn = nn.Module()
            blk.attn.qkv = nn.Linear(spec.dim_in, 3 * spec.dim_out)
            blk.attn.proj = nn.Linear(spec.dim_out, spec.dim_out)
            blk.norm2 = nn.LayerNorm(spec.dim_out, eps=cfg.ln_eps)
            blk.mlp = nn.Module()
            hidden = int(spec.dim_out * cfg.mlp_ratio)
            blk.mlp.layers = nn.ModuleList([nn.Linear(spec.dim_out, hidden), nn.Linear(hidden, spec.dim_out)])
            if spec.dim_in != spec.dim_out:
                blk.proj = nn.Linear(spec.dim_in, spec.dim_out)
            self.blocks.append(blk)

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        sd = dict(self.named_parameters())
        return hiera_forward(sd, x, self.cfg)


# HF port parameter names -> upstream names (used only by the cross-check test).
def hf_key_to_upstream(key: str) -> str:
    key = key.replace("patch_embed.projection.", "patch_embed.proj.")
    key = key.replace(".layer_norm1.", ".norm1.").replace(".layer_norm2.", ".norm2.")
    key = key.replace(".mlp.proj_in.", ".mlp.layers.0.").replace(".mlp.proj_out.", ".mlp.layers.1.")
    return key
