"""Drop-in replacement for the reference ``models/spegnet.py::SPEGNet`` nn.Module.

Same constructor (``SPEGNet(config)``, models/spegnet.py:90-98), same state-dict key names
(``encoder.encoder.<sam2 Hiera names>``, ``fusion.*``, ``context.*``, ``edge_detector.*``,
``decoder.*``), same ``forward(x[B,3,S,S] fp32) -> {'predictions': [p1,p2,p3], 'edge', 'features'}``
contract with fp32 logits (models/spegnet.py:137-206), same ``ValueError``s for malformed input
(models/feature_encoding.py:230-233).  All arithmetic of the forward runs in the hand-written sm_100a
kernels of ``libspegnet_b200.so`` (bf16 operands, fp32 accumulation, fp32 residual stream); PyTorch only
owns memory, streams and the one-time weight repack.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import math
import os
from collections.abc import Mapping
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .schema import (ASPP_DILATIONS, BN_EPS, LN_EPS, VARIANT_CHANNELS, TrunkSpec, head_entries, trunk_blocks,
                     trunk_param_shapes)


class _Node(nn.Module):
    """Anonymous container: only exists to reproduce the reference's dotted parameter names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("container module, not callable")


def _attach(root: nn.Module, name: str, tensor: torch.Tensor, kind: str) -> None:
    parts = name.split(".")
    mod = root
    for part in parts[:-1]:
        child = mod._modules.get(part)
        if child is None:
            child = _Node()
            mod.add_module(part, child)
        mod = child
    if kind == "param":
        mod.register_parameter(parts[-1], nn.Parameter(tensor))
    else:
        mod.register_buffer(parts[-1], tensor)


def _default_init(name: str, shape: Tuple[int, ...]) -> torch.Tensor:
    leaf = name.rsplit(".", 1)[-1]
    is_norm = any(t in name for t in (".norm1.", ".norm2.", ".bn.", ".bn1.", ".bn2.")) or _is_seq_bn(name)
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if leaf == "running_var" or (is_norm and leaf == "weight"):
        return torch.ones(shape)
    if leaf in ("running_mean", "bias") or "pos_embed" in name:
        return torch.zeros(shape)
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    return torch.randn(shape) * (1.0 / math.sqrt(max(fan_in, 1)))


def _is_seq_bn(name: str) -> bool:
    # BatchNorms that live inside nn.Sequential in the reference (numeric names)
    return any(name.startswith(p) for p in ("context.reduce.1.", "context.global_branch.2.", "context.fusion.1.",
                                            "context.expand.1.")) or (
        name.startswith("context.branches.") and name.split(".")[3] == "1")


def up2_operators(n: int = 8) -> torch.Tensor:
    """R[cls, a, t, d]: coefficient of low-resolution sample i + d - 1 in the bilinear x2 (align_corners=False) value
    at high-resolution position 2i + a + t - 1, for i in the first (cls 0) / an interior (1) / the last (2) row or
    column; positions outside the upsampled grid (the conv's zero padding) and samples outside the input contribute
    0.  Read off ATen's own operator (F.interpolate of unit impulses), so the clamp at the border is ATen's."""
    U = F.interpolate(torch.eye(n, dtype=torch.float64)[None], scale_factor=2, mode="linear", align_corners=False)[0].t()
    R = torch.zeros(3, 2, 3, 3, dtype=torch.float64)
    for cls, i in enumerate((0, n // 2, n - 1)):
        for a in range(2):
            for t in range(3):
                h = 2 * i + a + t - 1
                for d in range(3):
                    j = i + d - 1
                    if 0 <= h < 2 * n and 0 <= j < n:
                        R[cls, a, t, d] = U[h, j]
    return R


def up2_phase_weights(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fold `conv3x3(pad=1) o bilinear_x2` (models/object_detection.py:219,230) into convolutions on the low-resolution
    grid.  w [Cout,Cin,3,3] -> (main [3*4*Cout, 9*Cin], delta_left [4*Cout, 9*Cin], delta_right [4*Cout, 9*Cin]):
    main[cls*4*Cout + (a*2+b)*Cout + co, (dy*3+dx)*Cin + ci] for row class cls with interior column weights;
    delta_side[(a*2+b)*Cout + co, (cls*3+dy)*Cin + ci] = what the first / last image column adds on top (it only involves
    the border column itself).  See spg_conv3x3_up2_h16."""
    R = up2_operators().to(w.device, torch.float64)
    w64 = w.to(torch.float64)
    Co, Ci = w.shape[:2]
    main = torch.cat([torch.einsum("oiyx,ayd,bxe->abodei", w64, R[c], R[1]).reshape(4 * Co, 9 * Ci) for c in range(3)])
    deltas = []
    for side in (0, 2):
        dcol = (R[side] - R[1])[:, :, 1]  # [b, tx]: the other two offsets are unchanged (asserted in tests)
        deltas.append(torch.cat([torch.einsum("oiyx,ayd,bx->abodi", w64, R[c], dcol).reshape(4 * Co, 3 * Ci)
                                 for c in range(3)], dim=1))
    return main.to(w.dtype), deltas[0].to(w.dtype), deltas[1].to(w.dtype)


class LazyFeatures(Mapping):
    """``outputs['features']``: the reference returns fp32 NCHW tensors that no call site reads
    (SURVEY.md 3.3).  The kernels keep them as bf16 NHWC; this mapping converts on first access."""

    _KEYS = ("context", "fused", "edge_features")

    def __init__(self, raw: Dict[str, torch.Tensor]):
        self._raw = raw
        self._done: Dict[str, torch.Tensor] = {}

    def __getitem__(self, key: str) -> torch.Tensor:
        if key not in self._raw:
            raise KeyError(key)
        if key not in self._done:
            t = self._raw[key]  # [B,H,W,C] bf16
            B, H, W, C = t.shape
            out = torch.empty(B, C, H, W, dtype=torch.float32, device=t.device)
            ops.nhwc_to_nchw_f32(t, out, B, H * W, C)
            self._done[key] = out
        return self._done[key]

    def __iter__(self) -> Iterator[str]:
        return iter(self._KEYS)

    def __len__(self) -> int:
        return len(self._KEYS)


class _Workspace:
    """Device buffers for one (batch, resolution); reused across forwards on the same stream."""

    def __init__(self, B: int, S: int, dev: torch.device, spec: TrunkSpec, h16: torch.dtype, ln_records: bool = False,
                 blocks=()):
        bf, f32 = h16, torch.float32
        d = spec.dims
        G = S // 4
        T = [B * (G >> s) ** 2 for s in range(4)]  # tokens per stage
        e = lambda n, dt: torch.empty(n, dtype=dt, device=dev)  # noqa: E731
        self.cols = e(T[0] * 168, bf).view(T[0], 168)
        self.x = [e(T[s] * d[s], f32).view(T[s], d[s]) for s in range(4)]
        self.proj = e(max(T[s] * d[s + 1] for s in range(3)), f32)
        self.y = e(max(T[s] * d[s] for s in range(4)), bf)
        self.qkv = e(max(max(T[s] * 3 * d[s] for s in range(4)), max(T[s] * 3 * d[s + 1] for s in range(3))), bf)
        self.att = e(max(T[s] * d[s] for s in range(4)), bf)
        self.hid = e(max(T[s] * 4 * d[s] for s in range(4)), bf)
        self.f = [None] + [e(T[s] * d[s], bf).view(T[s], d[s]) for s in range(1, 4)]
        h = S // 8
        self.g = [e(T[s] * 512, f32).view(T[s], 512) for s in range(1, 4)]
        self.rs512 = e(B * h * 512, f32)
        self.gate = e(B * 512, f32)
        self.r128 = e(T[1] * 128, bf).view(T[1], 128)
        self.rs128 = e(B * h * 128, f32)
        self.gvec = e(B * 128, f32)
        self.y128 = e(T[1] * 128, bf).view(T[1], 128)
        self.u1 = e(B * 4 * h * h * 320, bf).view(B, 2 * h, 2 * h, 320)
        self.d1a = e(B * 4 * h * h * 256, bf).view(B, 2 * h, 2 * h, 256)
        self.d1 = e(B * 4 * h * h * 256, bf).view(B, 2 * h, 2 * h, 256)
        self.u2 = e(B * 16 * h * h * 320, bf).view(B, 4 * h, 4 * h, 320)
        self.d2a = e(B * 16 * h * h * 128, bf).view(B, 4 * h, 4 * h, 128)
        self.d2 = e(B * 16 * h * h * 128, bf).view(B, 4 * h, 4 * h, 128)
        # stage 3 runs the upsample fused into its first conv: border-column operand + corrections instead of a
        # materialised [B, 8h, 8h, 128] map
        self.bord = e(2 * B * 4 * h * 9 * 128, bf).view(2, B * 4 * h, 9 * 128)
        self.corr = e(2 * B * 4 * h * 256, f32).view(2, B * 4 * h, 256)
        self.d3a = e(B * 64 * h * h * 64, bf).view(B, 8 * h, 8 * h, 64)
        # LayerNorm row records {c, P, (sum, sum sq) x P} of the residual stream, ping-pong (see spg_epilogue_t)
        self.rec = [e(T[0] * 32, f32).view(T[0], 32) for _ in range(2)] if ln_records else None
        # window attention on grids that do not tile into the block's windows (S not a multiple of 256): zero-initialised
        # padded norm1 outputs per (padded edge, channels) -- only their valid region is ever rewritten -- and the padded
        # qkv / attention-output scratch
        self.ypad: Dict[Tuple[int, int], torch.Tensor] = {}
        qp = ap = 0
        Hs = G
        for b in blocks:
            if b.window and Hs % b.window:
                Hp = -(-Hs // b.window) * b.window
                if (Hp, b.dim_in) not in self.ypad:
                    self.ypad[(Hp, b.dim_in)] = torch.zeros(B, Hp, Hp, b.dim_in, dtype=bf, device=dev)
                Hop = Hp // 2 if b.q_pool else Hp
                qp = max(qp, B * Hp * Hp * 3 * b.dim_out)
                ap = max(ap, B * Hop * Hop * b.dim_out)
            if b.q_pool:
                Hs //= 2
        self.qkvp = e(qp, bf) if qp else None
        self.attp = e(ap, bf) if ap else None
        self.pos: Optional[torch.Tensor] = None  # [G*G, 144] fp32, set by the model (input independent)


class SPEGNet(nn.Module):
    """B200-native SPEGNet (inference).  See module docstring for the contract."""

    def __init__(self, config: Dict, compute_dtype: Optional[torch.dtype] = None, cuda_graph_max_batch: int = 8):
        super().__init__()
        # Batches up to this size replay a captured CUDA graph of the ~370-launch forward (launch overhead
        # otherwise dominates batch-1 latency); 0 disables.  Larger batches are GPU-bound and launch eagerly.
        self.cuda_graph_max_batch = int(os.environ.get("SPEGNET_B200_GRAPH_MAX_BATCH", cuda_graph_max_batch))
        self._graphs: Dict[Tuple[int, int, str], tuple] = {}
        # 16-bit storage type of activations / weights inside the kernels.  fp16 (default) and bf16 run the
        # same tcgen05 kind::f16 MMAs at the same rate; fp16's 10-bit mantissa is what meets the 1e-2 mask
        # parity bar on spread logits (DESIGN.md "Numerics"), bf16 is kept for range-critical checkpoints.
        if compute_dtype is None:
            compute_dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[
                os.environ.get("SPEGNET_B200_DTYPE", "fp16").lower()]
        if compute_dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("compute_dtype must be torch.float16 or torch.bfloat16")
        self.compute_dtype = compute_dtype
        enc = config["encoder"]
        # kept for parity with the reference constructor (models/spegnet.py:94-98); the SAM2 checkpoint is
        # not read here -- a SPEGNet checkpoint overwrites every trunk weight (engine/predictor.py:277-278).
        self.model_cfg = enc["config_path"]
        self.checkpoint_path = enc["checkpoint_path"]
        variant = enc.get("variant", "large")
        if variant not in VARIANT_CHANNELS:
            raise ValueError(f"Invalid variant. Choose from: {list(VARIANT_CHANNELS.keys())}")
        if variant != "large":
            raise NotImplementedError(f"spegnet_b200 implements the Hiera-'large' trunk only (got {variant!r})")
        self.variant = variant
        self.spec = TrunkSpec()
        self.blocks = trunk_blocks(self.spec)
        self.in_channels_list = VARIANT_CHANNELS[variant][1:4]
        for name, shape in trunk_param_shapes(self.spec).items():
            full = "encoder.encoder." + name
            _attach(self, full, _default_init(full, shape), "param")
        for name, (shape, kind) in head_entries(tuple(self.in_channels_list)).items():
            _attach(self, name, _default_init(name, shape), "param" if kind == "param" else "buffer")
        # LayerNorm folded into the GEMMs around it (`_trunk_ln_folded`): removes the 96 LayerNorm launches and 10 GB of
        # traffic per batch-64 step, but measured 972 vs 981 img/s at batch 64 (the residual GEMMs pay for the second
        # store and the consumers for reading the row records) and 3.00 vs 3.09 ms at batch 1, and the row statistics
        # then depend on the n-tiling, i.e. on the batch size, in the last bit.  Off by default (SPG_LN_FUSE=1 enables).
        # LayerNorm applied by the GEMM that produces its input (spg_epilogue_t.ln_apply_*; the CTAs holding a row block
        # form a cluster and exchange row statistics over distributed shared memory): widths 144 / 288 / 576.
        #   1 (default): the second MLP layer (K = 4 d: the two-pass epilogue hides behind the main loop) also emits the
        #      next block's norm1 -- measured faster than GEMM + LayerNorm kernel by 18 / 38 / 37 us per launch in stages
        #      3 / 2 / 1 (tools/ln_bench.py);
        #   2: additionally the attention projection and the patch embedding (K = d: the epilogue is the critical path
        #      there and the fused form measured 10-12 us SLOWER than the two kernels): no LayerNorm launch in stages 1-3;
        #   0: every LayerNorm is its own kernel.
        self.ln_apply = int(os.environ.get("SPG_LN_APPLY", "1"))
        # below this many rows the same LayerNorm runs as its own (bit-identical) kernel: spg_layernorm_matched_f32_h16
        self.ln_apply_min_rows = int(os.environ.get("SPG_LN_APPLY_MIN_ROWS", "4096"))
        self.ln_apply_widths = tuple(int(v) for v in os.environ.get("SPG_LN_APPLY_WIDTHS", "144,288,576").replace("+", ",").split(",") if v)
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._ln_fuse = os.environ.get("SPG_LN_FUSE", "0") != "0"
        self._debug_taps: Optional[Dict[str, torch.Tensor]] = None  # tests: stream snapshot after every block
        self._workspaces: Dict[Tuple[int, int, str], _Workspace] = {}
        self.eval()

    @property
    def ln_fuse(self) -> bool:
        return self._ln_fuse

    @ln_fuse.setter
    def ln_fuse(self, on: bool) -> None:
        if bool(on) != self._ln_fuse:  # the folded weights W diag(gamma) are only packed when the switch is on
            self._ln_fuse = bool(on)
            self._packed = None
            self._workspaces = {}
            self._graphs = {}

    # ------------------------------------------------------------------ nn.Module protocol hooks
    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        self._workspaces = {}
        self._graphs = {}
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self._packed = None
        self._graphs = {}
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def repack(self) -> None:
        """Call after modifying parameters in place (the bf16 / BN-folded copies are cached)."""
        self._packed = None
        self._graphs = {}

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("spegnet_b200.SPEGNet is inference only (BatchNorm uses running statistics)")
        return super().train(False)

    # ------------------------------------------------------------------ one-time weight repack
    @torch.no_grad()
    def _pack(self) -> Dict[str, torch.Tensor]:
        """One-time weight repack: BatchNorm folding, [N,K] 16-bit Linear / tap-major conv layouts, per-scale split of
        the fusion conv, the phase-folded stage-3 conv.  Runs on the HOST copy of the state dict (plain memcpys down
        and up, no ATen kernels on the device): the first kernels a process launches on the GPU are the library's own."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("spegnet_b200.SPEGNet runs on a CUDA (B200) device only; call .to('cuda') first")
        return {k: v.to(dev) for k, v in self._pack_host().items()}

    @torch.no_grad()
    def _pack_host(self) -> Dict[str, torch.Tensor]:
        """The packed weights as host tensors (everything `_pack` uploads; also what the CPU tests inspect)."""
        sd = {k: v.detach().to("cpu") for k, v in self.state_dict().items()}
        bf = self.compute_dtype
        W: Dict[str, torch.Tensor] = {}
        f32 = lambda t: t.to(torch.float32).contiguous()  # noqa: E731
        t = "encoder.encoder."
        # im2col column order of spg_patchify_7x7s4: k = (ky*3 + c)*8 + kx, slot kx = 7 zero -> K = 168
        pe = f32(sd[t + "patch_embed.proj.weight"]).permute(0, 2, 1, 3)  # [144, ky, c, kx]
        W["pe.w"] = F.pad(pe, (0, 1)).reshape(self.spec.embed_dim, 168).to(bf).contiguous()
        W["pe.b"] = f32(sd[t + "patch_embed.proj.bias"])
        W["pos_embed"] = f32(sd[t + "pos_embed"])
        W["pos_embed_window"] = f32(sd[t + "pos_embed_window"])
        for b in self.blocks:
            src, dst = f"{t}blocks.{b.index}.", f"b{b.index}."
            for a, c in (("norm1", "n1"), ("norm2", "n2")):
                W[dst + c + ".w"] = f32(sd[src + a + ".weight"])
                W[dst + c + ".b"] = f32(sd[src + a + ".bias"])
            for a, c in (("attn.qkv", "qkv"), ("attn.proj", "ap"), ("mlp.layers.0", "fc1"), ("mlp.layers.1", "fc2"),
                         ("proj", "proj")):
                if src + a + ".weight" in sd:
                    W[dst + c + ".w"] = f32(sd[src + a + ".weight"]).to(bf).contiguous()
                    W[dst + c + ".b"] = f32(sd[src + a + ".bias"])
            # LayerNorm folded into its consumers: W' = W diag(gamma) (16 bit), column sums of the ROUNDED W' (they
            # multiply the row mean that the 16-bit operand still carries), bias' = b + W beta in fp32
            for c, n in (("qkv", "n1"), ("proj", "n1"), ("fc1", "n2")):
                if self.ln_fuse and dst + c + ".w" in W:
                    wf = f32(sd[src + {"qkv": "attn.qkv", "proj": "proj", "fc1": "mlp.layers.0"}[c] + ".weight"])
                    wl = (wf * W[dst + n + ".w"][None, :]).to(bf).contiguous()
                    W[dst + c + ".wl"] = wl
                    W[dst + c + ".cw"] = wl.float().sum(dim=1).contiguous()
                    W[dst + c + ".bl"] = (W[dst + c + ".b"] + wf @ W[dst + n + ".b"]).contiguous()

        def bn_fold(prefix: str):
            g, be = f32(sd[prefix + "weight"]), f32(sd[prefix + "bias"])
            mu, var = f32(sd[prefix + "running_mean"]), f32(sd[prefix + "running_var"])
            s = g / torch.sqrt(var + BN_EPS)
            return s, be - mu * s

        # CFI fusion: split the 2016-column 1x1 conv per source scale, fold the BN scale into the rows
        s, sh = bn_fold("fusion.bn.")
        wf = f32(sd["fusion.conv1x1.weight"]).reshape(512, -1) * s[:, None]
        c2, c3, c4 = self.in_channels_list
        W["fu.w2"] = wf[:, :c2].to(bf).contiguous()
        W["fu.w3"] = wf[:, c2:c2 + c3].to(bf).contiguous()
        W["fu.w4"] = wf[:, c2 + c3:].to(bf).contiguous()
        W["fu.b"] = sh.contiguous()
        W["se.w1"] = f32(sd["fusion.se_block.fc.0.weight"])
        W["se.w2"] = f32(sd["fusion.se_block.fc.2.weight"])
        # e-ASPP
        s, sh = bn_fold("context.reduce.1.")
        W["red.w"] = (f32(sd["context.reduce.0.weight"]).reshape(128, 512) * s[:, None]).to(bf).contiguous()
        W["red.b"] = sh.contiguous()
        dw, dwb = [], []
        for i in range(4):
            s, sh = bn_fold(f"context.branches.{i}.1.")
            w = f32(sd[f"context.branches.{i}.0.weight"]).reshape(128, 9) * s[:, None]  # [ch, tap]
            dw.append(w.t().contiguous())  # [tap, ch]
            dwb.append(sh)
        W["aspp.dw"] = torch.stack(dw).contiguous()       # [4, 9, 128]
        W["aspp.dwb"] = torch.stack(dwb).contiguous()     # [4, 128]
        s, sh = bn_fold("context.global_branch.2.")
        W["glob.w"] = (f32(sd["context.global_branch.1.weight"]).reshape(128, 128) * s[:, None]).contiguous()
        W["glob.b"] = sh.contiguous()
        s, sh = bn_fold("context.fusion.1.")
        W["aspp.wf"] = (f32(sd["context.fusion.0.weight"]).reshape(128, 5) * s[:, None]).contiguous()
        W["aspp.wfb"] = sh.contiguous()
        s, sh = bn_fold("context.expand.1.")
        W["exp.w"] = (f32(sd["context.expand.0.weight"]).reshape(256, 128) * s[:, None]).to(bf).contiguous()
        W["exp.b"] = sh.contiguous()

        def conv3(wkey: str, bkey: Optional[str], bnp: str, out: str, up2: bool = False):
            s, sh = bn_fold(bnp)
            w = f32(sd[wkey])  # [Cout, Cin, 3, 3]
            w = w * s[:, None, None, None]
            bias = (sh + (f32(sd[bkey]) * s if bkey else 0.0)).contiguous()
            if up2:  # the conv consumes a bilinear x2 upsample: fold it in (4 phases on the low-resolution grid)
                main, dl, dr = up2_phase_weights(w)
                W[out + ".wp"] = main.to(bf).contiguous()
                W[out + ".dwl"] = dl.to(bf).contiguous()
                W[out + ".dwr"] = dr.to(bf).contiguous()
                W[out + ".bp"] = bias.repeat(4).contiguous()
                return
            W[out + ".w"] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(bf).contiguous()  # [Cout, (ky,kx,ci)]
            W[out + ".b"] = bias

        conv3("edge_detector.conv1.weight", None, "edge_detector.bn1.", "edge")
        W["edge.hw"] = f32(sd["edge_detector.edge_conv.weight"]).reshape(-1)
        W["edge.hb"] = f32(sd["edge_detector.edge_conv.bias"])
        for i in range(3):
            p = f"decoder.decoder_blocks.{i}."
            conv3(p + "conv1.weight", p + "conv1.bias", p + "bn1.", f"dec{i}a", up2=(i == 2))
            conv3(p + "conv2.weight", p + "conv2.bias", p + "bn2.", f"dec{i}b")
            W[f"head{i}.w"] = f32(sd[f"decoder.pred_heads.{i}.weight"]).reshape(-1)
            W[f"head{i}.b"] = f32(sd[f"decoder.pred_heads.{i}.bias"])
        # scalar head biases as python floats (kernel arguments), read once here rather than per forward
        self._head_b = {k: float(W[k].item()) for k in ("edge.hb", "head0.b", "head1.b", "head2.b")}
        if bf == torch.float16:
            # fp16 storage: a folded weight beyond the fp16 range would become inf here (activations saturate in the
            # kernels, half16.cuh); refuse it rather than produce NaN masks
            for k, v in W.items():
                if v.dtype == torch.float16 and not bool(torch.isfinite(v).all()):
                    raise ValueError(f"packed weight {k!r} overflows fp16 (|w| > 65504): construct the model with "
                                     "compute_dtype=torch.bfloat16 for this checkpoint")
        self._pos_src = (W.pop("pos_embed"), W.pop("pos_embed_window"))  # host copies: interpolated per resolution
        return W

    def _pos_map(self, W: Dict[str, torch.Tensor], G: int) -> torch.Tensor:
        """bicubic(pos_embed 7x7 -> GxG) + tiled window embedding, [G*G, 144] fp32 (HF:modeling_sam2.py:623-629).
        Input independent: computed once per resolution at plan time, then fused as a residual of the
        patch-embedding GEMM."""
        pos, win = self._pos_src
        bg = F.interpolate(pos, size=(G, G), mode="bicubic")
        tiled = win.repeat(1, 1, G // win.shape[-2], G // win.shape[-1])
        return (bg + tiled).permute(0, 2, 3, 1).reshape(G * G, -1).contiguous()

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Dict[str, object]:
        if x.dim() != 4:
            raise ValueError(f"Expected 4D input (B,C,H,W), got {x.dim()}D")
        if any(s % 32 != 0 for s in x.shape[-2:]):
            raise ValueError("Input spatial dims must be divisible by 32")
        B, Cin, S, S2 = x.shape
        if Cin != 3 or S != S2:
            raise ValueError(f"Expected square RGB input [B,3,S,S], got {tuple(x.shape)}")
        # any S % 32 == 0 (352, 384, ... as well as 512 / 1024): token grids that do not tile into the block's windows
        # run through zero-padded grids around the attention (HF:modeling_sam2.py:395-399, `_trunk`), and the head's
        # convolutions take a ragged last tile column (spg_conv3x3_h16)
        if not x.is_cuda:
            raise RuntimeError("spegnet_b200.SPEGNet needs a CUDA tensor on a B200; there is no CPU fallback")
        if self._packed is None:
            self._packed = self._pack()
        x = x.contiguous().float()
        with torch.cuda.device(x.device):  # kernels, their attributes and the stream belong to the input's device
            # programmatic dependent launch pays in the latency regime only (DESIGN.md "Measurement"); per-call flag
            ops.set_pdl(B <= max(self.cuda_graph_max_batch, 8))
            if 0 < B <= self.cuda_graph_max_batch and self._debug_taps is None and not torch.cuda.is_current_stream_capturing():
                return self._forward_graphed(x, B, S)
            return self._forward_eager(x, B, S)

    def _new_workspace(self, x: torch.Tensor, B: int, S: int) -> _Workspace:
        ws = _Workspace(B, S, x.device, self.spec, self.compute_dtype, ln_records=self.ln_fuse, blocks=self.blocks)
        ws.pos = self._pos_map(self._packed, S // 4).to(x.device)
        return ws

    def _forward_eager(self, x: torch.Tensor, B: int, S: int, ws: Optional[_Workspace] = None) -> Dict[str, object]:
        W = self._packed
        if ws is None:
            key = (B, S, str(x.device))
            ws = self._workspaces.get(key)
            if ws is None:
                self._workspaces = {}  # eager workspaces are large: one live shape at a time
                ws = self._new_workspace(x, B, S)
                self._workspaces[key] = ws
        self._trunk(W, ws, x, B, S)
        return self._head(W, ws, B, S)

    def _forward_graphed(self, x: torch.Tensor, B: int, S: int) -> Dict[str, object]:
        """Small batches: capture the whole launch sequence once per (B, S) into a CUDA graph with static input /
        output buffers, then replay.  Callers still get fresh tensors (cloned from the static outputs)."""
        key = (B, S, str(x.device))
        entry = self._graphs.get(key)
        if entry is None:
            with torch.inference_mode(False):
                # a normal (non-inference) tensor: later calls may arrive under plain no_grad and copy into it
                static_x = torch.empty(x.shape, dtype=x.dtype, device=x.device)
            static_x.copy_(x)
            ws = self._new_workspace(x, B, S)  # owned by the graph entry: its addresses are baked into the graph
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: smem attributes, driver entry points
                self._forward_eager(static_x, B, S, ws)
                self._forward_eager(static_x, B, S, ws)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward_eager(static_x, B, S, ws)
            entry = (graph, static_x, static_out, ws)
            self._graphs[key] = entry
        graph, static_x, static_out, _ = entry
        static_x.copy_(x)
        graph.replay()
        raw = static_out["features"]._raw
        return {"predictions": [p.clone() for p in static_out["predictions"]], "edge": static_out["edge"].clone(),
                "features": LazyFeatures({k: v.clone() for k, v in raw.items()})}

    def _trunk(self, W, ws: _Workspace, x: torch.Tensor, B: int, S: int) -> None:
        if self.ln_fuse and self._debug_taps is None and ws.rec is not None and not ws.ypad:
            return self._trunk_ln_folded(W, ws, x, B, S)
        G = S // 4
        blocks = self.blocks
        ends = self.spec.stage_ends
        fused_widths = self.ln_apply_widths

        def ln_after(width: int, gamma_key: str, rows: int, long_k: bool):
            """How the LayerNorm `gamma_key` over `width` channels of a residual GEMM's output is computed: returns
            (ln_apply tuple or None, y, kernel).  With a tuple the producing GEMM stores y itself; otherwise `kernel` has
            to be launched on the fp32 stream afterwards.  For the widths the producers support, the separate kernel is
            the bit-identical `layernorm_matched`, so the choice (large batches: fused; latency regime: separate, the
            two-pass epilogue and the cluster launch cost ~5 us per GEMM there) never changes a result."""
            y = ws.y[: rows * width].view(rows, width)
            if width in fused_widths and self.ln_apply >= (1 if long_k else 2):
                if rows >= self.ln_apply_min_rows:
                    return (W[gamma_key + ".w"], W[gamma_key + ".b"], y, LN_EPS), y, None
                return None, y, ops.layernorm_matched
            return None, y, ops.layernorm

        ops.patchify(x, ws.cols)
        # the patch embedding (+ positional embedding as a broadcast residual) feeds block 0's norm1
        ap, y, ln1 = ln_after(blocks[0].dim_in, "b0.n1", B * G * G, False)
        ops.linear(ws.cols, W["pe.w"], ws.x[0], bias=W["pe.b"], residual=ws.pos, res_rows=G * G, ln_apply=ap)
        H = G
        cur = ws.x[0]
        if self._debug_taps is not None:
            self._debug_taps["embed"] = cur.view(B, G, G, -1).clone()
        for bi, b in enumerate(blocks):
            p = f"b{b.index}."
            M = B * H * H
            y = ws.y[: M * b.dim_in].view(M, b.dim_in)
            if ln1 is not None:
                ln1(cur, W[p + "n1.w"], W[p + "n1.b"], y, LN_EPS)
            if b.dim_in != b.dim_out:
                if not b.q_pool:
                    raise NotImplementedError("channel change without query pooling does not occur in Hiera-L")
                Ho = H // 2
                Mo = B * Ho * Ho
                nxt = ws.x[b.stage]
                proj = ws.proj[: M * b.dim_out].view(M, b.dim_out)
                ops.linear(y, W[p + "proj.w"], proj, bias=W[p + "proj.b"])
                ops.maxpool2x2(proj, nxt, B, H, H, b.dim_out)  # pooled shortcut lands in the new stream
            else:
                Ho, Mo, nxt = H, M, cur
            att = ws.att[: Mo * b.dim_out].view(Mo, b.dim_out)
            if b.window and H % b.window:
                # the grid does not tile into this block's windows: norm1 output -> zero-padded grid -> qkv (the padded
                # tokens come out as the bias and act as keys) -> window attention on the padded grid -> crop
                # (window_partition / window_unpartition, HF:modeling_sam2.py:395-399, 435-437, 514-521)
                Hp = -(-H // b.window) * b.window
                Hop = Hp // 2 if b.q_pool else Hp
                yp = ws.ypad[(Hp, b.dim_in)]
                ops.copy_grid(y.view(B, H, H, b.dim_in), yp, H, H)
                qkv = ws.qkvp[: B * Hp * Hp * 3 * b.dim_out].view(B * Hp * Hp, 3 * b.dim_out)
                ops.linear(yp.view(B * Hp * Hp, b.dim_in), W[p + "qkv.w"], qkv, bias=W[p + "qkv.b"])
                attp = ws.attp[: B * Hop * Hop * b.dim_out].view(B, Hop, Hop, b.dim_out)
                ops.window_attention(qkv, attp.view(-1, b.dim_out), B, Hp, Hp, b.dim_out, b.heads, b.window, b.q_pool)
                ops.copy_grid(attp, att.view(B, Ho, Ho, b.dim_out), Ho, Ho)
            else:
                qkv = ws.qkv[: M * 3 * b.dim_out].view(M, 3 * b.dim_out)
                ops.linear(y, W[p + "qkv.w"], qkv, bias=W[p + "qkv.b"])
                ops.window_attention(qkv, att, B, H, H, b.dim_out, b.heads, b.window, b.q_pool)
            # attention projection + residual; its output feeds norm2
            ap, z, ln2 = ln_after(b.dim_out, p + "n2", Mo, False)
            ops.linear(att, W[p + "ap.w"], nxt, bias=W[p + "ap.b"], residual=nxt, ln_apply=ap)
            if ln2 is not None:
                ln2(nxt, W[p + "n2.w"], W[p + "n2.b"], z, LN_EPS)
            hid = ws.hid[: Mo * 4 * b.dim_out].view(Mo, 4 * b.dim_out)
            ops.linear(z, W[p + "fc1.w"], hid, bias=W[p + "fc1.b"], act=ops.ACT_GELU)
            # second MLP layer + residual; its output feeds the NEXT block's norm1 (nothing after the last block)
            ap, ln1 = None, None
            if bi + 1 < len(blocks):
                ap, _, ln1 = ln_after(b.dim_out, f"b{blocks[bi + 1].index}.n1", Mo, True)
            ops.linear(hid, W[p + "fc2.w"], nxt, bias=W[p + "fc2.b"], residual=nxt, ln_apply=ap)
            cur, H = nxt, Ho
            if self._debug_taps is not None:
                self._debug_taps[f"block{b.index}"] = cur.view(B, H, H, b.dim_out).clone()
            if b.index in ends and b.stage >= 1:
                ops.cast_h16(cur, ws.f[b.stage])

    def _trunk_ln_folded(self, W, ws: _Workspace, x: torch.Tensor, B: int, S: int) -> None:
        """The trunk with every LayerNorm folded into the GEMMs around it (no LayerNorm launches): each residual GEMM
        (patch embed + pos, attention proj, fc2) also emits the centred 16-bit copy of its output rows and their
        statistics; qkv / fc1 / the dim-change proj consume them (spg_epilogue_t, `ln_*` fields)."""
        G = S // 4
        ends = self.spec.stage_ends
        rec_cur, rec_oth = ws.rec
        ops.patchify(x, ws.cols)
        H = G
        cur = ws.x[0]
        xh = ws.y[: B * G * G * self.spec.embed_dim].view(B * G * G, self.spec.embed_dim)
        ops.linear(ws.cols, W["pe.w"], cur, bias=W["pe.b"], residual=ws.pos, res_rows=G * G,
                   ln_emit=(rec_cur[: B * G * G], None, xh))
        last = self.blocks[-1].index
        for b in self.blocks:
            p = f"b{b.index}."
            M = B * H * H
            r1 = rec_cur[:M]
            if b.dim_in != b.dim_out:
                if not b.q_pool:
                    raise NotImplementedError("channel change without query pooling does not occur in Hiera-L")
                Ho = H // 2
                Mo = B * Ho * Ho
                nxt = ws.x[b.stage]
                proj = ws.proj[: M * b.dim_out].view(M, b.dim_out)
                ops.linear(xh, W[p + "proj.wl"], proj, bias=W[p + "proj.bl"], ln_fold=(r1, W[p + "proj.cw"], b.dim_in, LN_EPS))
                ops.maxpool2x2(proj, nxt, B, H, H, b.dim_out)  # pooled shortcut lands in the new stream
                prev = None                                    # new stream: no earlier statistics of these rows
            else:
                Ho, Mo, nxt = H, M, cur
                prev = r1
            qkv = ws.qkv[: M * 3 * b.dim_out].view(M, 3 * b.dim_out)
            ops.linear(xh, W[p + "qkv.wl"], qkv, bias=W[p + "qkv.bl"], ln_fold=(r1, W[p + "qkv.cw"], b.dim_in, LN_EPS))
            att = ws.att[: Mo * b.dim_out].view(Mo, b.dim_out)
            ops.window_attention(qkv, att, B, H, H, b.dim_out, b.heads, b.window, b.q_pool)
            z = ws.y[: Mo * b.dim_out].view(Mo, b.dim_out)
            r2 = rec_oth[:Mo]
            ops.linear(att, W[p + "ap.w"], nxt, bias=W[p + "ap.b"], residual=nxt, ln_emit=(r2, prev, z))
            hid = ws.hid[: Mo * 4 * b.dim_out].view(Mo, 4 * b.dim_out)
            ops.linear(z, W[p + "fc1.wl"], hid, bias=W[p + "fc1.bl"], act=ops.ACT_GELU,
                       ln_fold=(r2, W[p + "fc1.cw"], b.dim_out, LN_EPS))
            if b.index != last:  # the next block's LayerNorm reads this output
                xh = ws.y[: Mo * b.dim_out].view(Mo, b.dim_out)
                ops.linear(hid, W[p + "fc2.w"], nxt, bias=W[p + "fc2.b"], residual=nxt, ln_emit=(rec_cur[:Mo], r2, xh))
            else:
                ops.linear(hid, W[p + "fc2.w"], nxt, bias=W[p + "fc2.b"], residual=nxt)
            cur, H = nxt, Ho
            if b.index in ends and b.stage >= 1:
                ops.cast_h16(cur, ws.f[b.stage])

    def _head(self, W, ws: _Workspace, B: int, S: int) -> Dict[str, object]:
        dev = ws.cols.device
        bf = self.compute_dtype
        h = S // 8
        T2 = B * h * h
        # fresh, caller-owned outputs (callers keep references across their per-image loops)
        fused = torch.empty(B, h, h, 512, dtype=bf, device=dev)
        ctx = torch.empty(B, h, h, 256, dtype=bf, device=dev)
        ef = torch.empty(B, h, h, 64, dtype=bf, device=dev)
        edge = torch.empty(B, 1, h, h, dtype=torch.float32, device=dev)
        preds = [torch.empty(B, 1, h << (i + 1), h << (i + 1), dtype=torch.float32, device=dev) for i in range(3)]

        # ---- CFI: per-scale 1x1 products, combine (+BN+ReLU), squeeze-excite
        for i in range(3):
            ops.linear(ws.f[i + 1], W[f"fu.w{i + 2}"], ws.g[i])
        ops.fusion_combine(ws.g[0], ws.g[1], ws.g[2], W["fu.b"], fused, ws.rs512, B, h, 512)
        ops.pooled_mlp(ws.rs512, h, h * h, W["se.w1"], None, 32, W["se.w2"], ws.gate, B, 512)
        ops.scale_channels(fused, ws.gate, B, h * h, 512)
        # ---- e-ASPP
        ops.linear(fused.view(T2, 512), W["red.w"], ws.r128, bias=W["red.b"], act=ops.ACT_RELU)
        ops.row_sums(ws.r128, ws.rs128, B, h, h, 128)
        ops.pooled_mlp(ws.rs128, h, h * h, W["glob.w"], W["glob.b"], 128, None, ws.gvec, B, 128)
        ops.easpp_branches(ws.r128, W["aspp.dw"], W["aspp.dwb"], ws.gvec, W["aspp.wf"], W["aspp.wfb"], ws.y128, B, h, h,
                           ASPP_DILATIONS)
        ops.linear(ws.y128, W["exp.w"], ctx.view(T2, 256), bias=W["exp.b"], act=ops.ACT_RELU)
        # ---- EFE: 3x3 conv + BN + ReLU, edge logits fused into the epilogue
        ops.conv3x3(ctx, W["edge.w"], ef.view(T2, 64), bias=W["edge.b"], act=ops.ACT_RELU, head_w=W["edge.hw"],
                    head_b=self._head_b["edge.hb"], head_out=edge)
        # ---- PED: three stages, prediction heads fused into the second conv of each stage
        ops.upsample_concat(ctx, ef, ws.u1)
        ops.conv3x3(ws.u1, W["dec0a.w"], ws.d1a.view(-1, 256), bias=W["dec0a.b"], act=ops.ACT_RELU)
        ops.conv3x3(ws.d1a, W["dec0b.w"], ws.d1.view(-1, 256), bias=W["dec0b.b"], act=ops.ACT_RELU,
                    head_w=W["head0.w"], head_b=self._head_b["head0.b"], head_out=preds[0])
        ops.upsample_concat(ws.d1, ef, ws.u2)
        ops.conv3x3(ws.u2, W["dec1a.w"], ws.d2a.view(-1, 128), bias=W["dec1a.b"], act=ops.ACT_RELU)
        ops.conv3x3(ws.d2a, W["dec1b.w"], ws.d2.view(-1, 128), bias=W["dec1b.b"], act=ops.ACT_RELU,
                    head_w=W["head1.w"], head_b=self._head_b["head1.b"], head_out=preds[1])
        # stage 3 (no edge branch): x2 upsample folded into conv1 -- border-column corrections, then one N=256 conv on
        # the 4h x 4h grid with a pixel-shuffle store
        ops.up2_border_gather(ws.d2, ws.bord)
        ops.linear(ws.bord[0], W["dec2a.dwl"], ws.corr[0])
        ops.linear(ws.bord[1], W["dec2a.dwr"], ws.corr[1])
        ops.conv3x3_up2(ws.d2, W["dec2a.wp"], ws.corr, W["dec2a.bp"], ws.d3a)
        # stage-3 features never reach HBM: only the fused head output is stored
        ops.conv3x3(ws.d3a, W["dec2b.w"], None, bias=W["dec2b.b"], act=ops.ACT_RELU, head_w=W["head2.w"],
                    head_b=self._head_b["head2.b"], head_out=preds[2])
        return {"predictions": preds, "edge": edge,
                "features": LazyFeatures({"context": ctx, "fused": fused, "edge_features": ef})}

    # ------------------------------------------------------------------ debugging / tests
    @torch.no_grad()
    def encoder_features(self, x: torch.Tensor) -> List[torch.Tensor]:
        """The encoder's stage 2-4 maps as fp32 NCHW (test hook; the forward never materialises them)."""
        B, _, S, _ = x.shape
        if self._packed is None:
            self._packed = self._pack()
        self._forward_eager(x.contiguous().float(), B, S)  # fills the eager workspace
        ws = self._workspaces[(B, S, str(x.device))]
        feats = []
        for s in range(1, 4):
            hs = (S // 4) >> s
            c = self.spec.dims[s]
            feats.append(ws.x[s].view(B, hs, hs, c).permute(0, 3, 1, 2).contiguous())
        return feats
