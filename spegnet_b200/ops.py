"""Tensor-level wrappers over the C-ABI (one Python function per ``spg_*`` entry point).

Each wrapper validates dtype / device / contiguity, takes raw ``data_ptr()``s and enqueues on torch's
current CUDA stream.  No arithmetic happens in Python or in PyTorch ops here.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, F32, Epilogue  # noqa: F401
from ._lib import H16 as OUT_H16


class _LaunchState(threading.local):
    """Host-side launch policy of the calling thread; handed to the library per call (spg_launch_t), never stored in it."""

    pdl = False      # programmatic dependent launch: pays in the latency regime (small batches), see SPEGNet.forward
    direction = 0    # traversal direction of the next direction-aware launch (alternates: DESIGN.md "Launch structure")


_state = _LaunchState()
_SNAKE = os.environ.get("SPG_SNAKE", "1") != "0"


def set_pdl(on: bool) -> None:
    """Request programmatic dependent launch for the calls this thread makes from now on (SPG_LAUNCH_PDL)."""
    _state.pdl = bool(on)


def _launch(flip: bool = False):
    """spg_launch_t for the next call: torch's current stream + this thread's flags.  `flip` alternates the traversal
    direction (GEMM / conv / LayerNorm / attention): each consumer starts on the rows its producer wrote last."""
    flags = _lib.LAUNCH_PDL if _state.pdl else 0
    if flip and _SNAKE:
        _state.direction ^= 1
    if flip and _state.direction:
        flags |= _lib.LAUNCH_REVERSE
    return C.byref(_lib.Launch(torch.cuda.current_stream().cuda_stream, flags))


H16 = "h16"  # marker: a 16-bit operand (fp16 or bf16; selects the library variant)


def _on_current_device(t: torch.Tensor) -> None:
    """Kernels, their per-device attributes and the stream all belong to the CURRENT CUDA device: an operand that lives
    on another GPU is an error, not something to launch on (wrap the call in `torch.cuda.device(t.device)`)."""
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise ValueError(f"operand is on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "call under `with torch.cuda.device(tensor.device):`")


def _lib_for(t: torch.Tensor):
    """(library, dtype name) for the 16-bit operand `t`."""
    _on_current_device(t)
    name = _lib.dtype_name(t.dtype)
    return _lib.load(name), name


def _default_lib(t: torch.Tensor):
    """(library, dtype name) for the entry points without a 16-bit operand (either variant serves them)."""
    _on_current_device(t)
    return _lib.load(), _lib.DEFAULT_DTYPE


def _ptr(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (spegnet_b200 has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is H16:
        if t.dtype not in (torch.float16, torch.bfloat16):
            raise ValueError(f"{name} must be fp16 or bf16, got {t.dtype}")
    elif dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _epilogue(out, bias, act, residual, res_rows, head_w, head_b, head_out, ln_fold=None, ln_emit=None,
              ln_apply=None) -> Epilogue:
    ep = Epilogue()
    ep.bias = _ptr(bias, torch.float32, "bias")
    ep.act = act
    ep.residual = _ptr(residual, torch.float32, "residual")
    ep.res_rows = res_rows
    if out is not None:
        if out.dtype not in (torch.bfloat16, torch.float16, torch.float32):
            raise ValueError("out must be fp16 / bf16 or fp32")
        ep.out = _ptr(out, None, "out")
        ep.out_dtype = F32 if out.dtype == torch.float32 else OUT_H16
    ep.head_w = _ptr(head_w, torch.float32, "head_w")
    ep.head_b = float(head_b)
    ep.head_out = _ptr(head_out, torch.float32, "head_out")
    if ln_fold is not None:  # (records [M,32] fp32, colsum(W') [N] fp32, channels, eps): see spg_epilogue_t
        rec, cw, cols, eps = ln_fold
        ep.ln_fold_rec = _ptr(rec, torch.float32, "ln_fold_rec")
        ep.ln_fold_cw = _ptr(cw, torch.float32, "ln_fold_cw")
        ep.ln_cols, ep.ln_eps = int(cols), float(eps)
    if ln_emit is not None:  # (records out [M,32] fp32, previous records or None, centred 16-bit copy [M,N])
        rec, prev, xc = ln_emit
        ep.ln_emit_rec = _ptr(rec, torch.float32, "ln_emit_rec")
        ep.ln_prev_rec = _ptr(prev, torch.float32, "ln_prev_rec")
        ep.ln_emit_out = _ptr(xc, H16, "ln_emit_out")
        ep.ln_cols = int(xc.shape[1])
    if ln_apply is not None:  # (gamma [N] fp32, beta [N] fp32, y [M,N] 16-bit, eps): the producer normalises its own rows
        gamma, beta, y, eps = ln_apply
        ep.ln_apply_gamma = _ptr(gamma, torch.float32, "ln_apply_gamma")
        ep.ln_apply_beta = _ptr(beta, torch.float32, "ln_apply_beta")
        ep.ln_apply_out = _ptr(y, H16, "ln_apply_out")
        ep.ln_eps = float(eps)
    return ep


def linear(a: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor], *, bias=None, act=ACT_NONE,
           residual=None, res_rows: int = 0, head_w=None, head_b: float = 0.0, head_out=None, ln_fold=None,
           ln_emit=None, ln_apply=None) -> None:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T); a, w bf16; see spg_linear_h16.  `ln_apply` = (gamma, beta, y, eps) makes a
    residual GEMM also store y = LayerNorm(out) (the next GEMMs' operand; no separate LayerNorm pass).  `ln_fold` /
    `ln_emit` are the older folded formulation (opt-in), see spg_epilogue_t."""
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"weight K {w.shape[1]} != activation K {K}")
    ep = _epilogue(out, bias, act, residual, res_rows, head_w, head_b, head_out, ln_fold, ln_emit, ln_apply)
    lib, dn = _lib_for(a)
    rc = lib.spg_linear_h16(_ptr(a, H16, "a"), _ptr(w, H16, "w"), M, N, K,
                                     C.byref(ep), _launch(True))
    _lib.check(rc, "spg_linear_h16", dn)


def conv3x3(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor], *, bias=None, act=ACT_NONE,
            head_w=None, head_b: float = 0.0, head_out=None) -> None:
    """x bf16 NHWC [B,H,W,Cin], w bf16 [Cout, 9*Cin] (tap-major); see spg_conv3x3_h16."""
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    if w.shape[1] != 9 * Cin:
        raise ValueError("conv weight must be [Cout, 9*Cin]")
    ep = _epilogue(out, bias, act, None, 0, head_w, head_b, head_out)
    lib, dn = _lib_for(x)
    rc = lib.spg_conv3x3_h16(_ptr(x, H16, "x"), _ptr(w, H16, "w"), B, H, W, Cin, Cout,
                                      C.byref(ep), _launch(True))
    _lib.check(rc, "spg_conv3x3_h16", dn)


def up2_border_gather(x: torch.Tensor, out: torch.Tensor) -> None:
    """x [B,H,W,C] h16 -> out [2, B*H, 9*C] h16 (left operand of the border-column correction GEMM of conv3x3_up2)."""
    B, H, W, Cc = x.shape
    if tuple(out.shape) != (2, B * H, 9 * Cc):
        raise ValueError(f"out must be [2, {B * H}, {9 * Cc}], got {tuple(out.shape)}")
    lib, dn = _lib_for(x)
    rc = lib.spg_up2_border_gather_h16(_ptr(x, H16, "x"), _ptr(out, H16, "out"), B, H, W, Cc, _launch())
    _lib.check(rc, "spg_up2_border_gather_h16", dn)


def conv3x3_up2(x: torch.Tensor, w_phase: torch.Tensor, corr: torch.Tensor, bias4: torch.Tensor, out: torch.Tensor) -> None:
    """out [B,2H,2W,Cout] = relu(conv3x3(bilinear_x2(x)) + bias) without materialising the upsampled map.
    x [B,H,W,Cin] h16, w_phase [3*4*Cout, 9*Cin] h16, corr [2, B*H, 4*Cout] fp32, bias4 [4*Cout] fp32."""
    B, H, W, Cin = x.shape
    Cout = w_phase.shape[0] // 12
    if tuple(w_phase.shape) != (12 * Cout, 9 * Cin):
        raise ValueError(f"w_phase must be [12*Cout, 9*Cin], got {tuple(w_phase.shape)}")
    if tuple(out.shape) != (B, 2 * H, 2 * W, Cout) or tuple(corr.shape) != (2, B * H, 4 * Cout):
        raise ValueError("out / corr shape does not match x and w_phase")
    lib, dn = _lib_for(x)
    rc = lib.spg_conv3x3_up2_h16(_ptr(x, H16, "x"), _ptr(w_phase, H16, "w_phase"), _ptr(corr, torch.float32, "corr"),
                                 B, H, W, Cin, Cout, _ptr(bias4, torch.float32, "bias4"), _ptr(out, H16, "out"), _launch(True))
    _lib.check(rc, "spg_conv3x3_up2_h16", dn)


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, y: torch.Tensor, eps: float) -> None:
    M, Cc = x.shape
    lib, dn = _lib_for(y)
    rc = lib.spg_layernorm_f32_h16(_ptr(x, torch.float32, "x"), _ptr(gamma, torch.float32, "gamma"),
                                            _ptr(beta, torch.float32, "beta"), _ptr(y, H16, "y"), M, Cc,
                                            eps, _launch(True))
    _lib.check(rc, "spg_layernorm_f32_h16", dn)


def layernorm_matched(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, y: torch.Tensor, eps: float) -> None:
    """LayerNorm bit-identical to `linear(..., ln_apply=...)` on the same rows (spg_layernorm_matched_f32_h16)."""
    M, Cc = x.shape
    lib, dn = _lib_for(y)
    rc = lib.spg_layernorm_matched_f32_h16(_ptr(x, torch.float32, "x"), _ptr(gamma, torch.float32, "gamma"),
                                           _ptr(beta, torch.float32, "beta"), _ptr(y, H16, "y"), M, Cc, eps, _launch(True))
    _lib.check(rc, "spg_layernorm_matched_f32_h16", dn)


def copy_grid(src: torch.Tensor, dst: torch.Tensor, H: int, W: int) -> None:
    """dst[:, :H, :W] = src[:, :H, :W] for h16 NHWC grids [B,Hs,Ws,C] / [B,Hd,Wd,C] (pad / crop of a token grid)."""
    B, Hs, Ws, Cc = src.shape
    Bd, Hd, Wd, Cd = dst.shape
    if B != Bd or Cc != Cd:
        raise ValueError(f"grids differ in batch / channels: {tuple(src.shape)} vs {tuple(dst.shape)}")
    lib, dn = _lib_for(src)
    rc = lib.spg_copy_grid_h16(_ptr(src, H16, "src"), Hs, Ws, _ptr(dst, H16, "dst"), Hd, Wd, B, H, W, Cc, _launch())
    _lib.check(rc, "spg_copy_grid_h16", dn)


def patchify(x: torch.Tensor, cols: torch.Tensor) -> None:
    B, _, S, _ = x.shape
    lib, dn = _lib_for(cols)
    rc = lib.spg_patchify_7x7s4(_ptr(x, torch.float32, "x"), _ptr(cols, H16, "cols"), B, S, _launch())
    _lib.check(rc, "spg_patchify_7x7s4", dn)


def maxpool2x2(x: torch.Tensor, y: torch.Tensor, B: int, H: int, W: int, Cc: int) -> None:
    lib, dn = _default_lib(x)
    rc = lib.spg_maxpool2x2_f32(_ptr(x, torch.float32, "x"), _ptr(y, torch.float32, "y"), B, H, W, Cc, _launch())
    _lib.check(rc, "spg_maxpool2x2_f32", dn)


def cast_h16(x: torch.Tensor, y: torch.Tensor) -> None:
    lib, dn = _lib_for(y)
    rc = lib.spg_cast_f32_h16(_ptr(x, torch.float32, "x"), _ptr(y, H16, "y"), x.numel(), _launch())
    _lib.check(rc, "spg_cast_f32_h16", dn)


def window_attention(qkv: torch.Tensor, out: torch.Tensor, B: int, H: int, W: int, D: int, heads: int, window: int,
                     q_pool: bool) -> None:
    lib, dn = _lib_for(qkv)
    rc = lib.spg_window_attention_h16(_ptr(qkv, H16, "qkv"), _ptr(out, H16, "out"), B,
                                               H, W, D, heads, window, int(q_pool), _launch(True))
    _lib.check(rc, "spg_window_attention_h16", dn)


def window_attention_tc(qkv: torch.Tensor, out: torch.Tensor, B: int, H: int, W: int, D: int, heads: int, window: int,
                        q_pool: bool) -> None:
    """The tcgen05 / TMEM attention kernels directly (16x16 windows or global, no query pooling); `window_attention`
    dispatches to them whenever they apply.  See spg_window_attention_tc_h16."""
    lib, dn = _lib_for(qkv)
    rc = lib.spg_window_attention_tc_h16(_ptr(qkv, H16, "qkv"), _ptr(out, H16, "out"), B, H, W, D, heads, window,
                                         int(q_pool), _launch(True))
    _lib.check(rc, "spg_window_attention_tc_h16", dn)


def upsample_concat(src0: torch.Tensor, src1: Optional[torch.Tensor], out: torch.Tensor) -> None:
    B, Ho, Wo, _ = out.shape
    _, h0, w0, c0 = src0.shape
    h1 = w1 = c1 = 0
    if src1 is not None:
        _, h1, w1, c1 = src1.shape
    lib, dn = _lib_for(src0)
    rc = lib.spg_upsample_concat_h16(_ptr(src0, H16, "src0"), h0, w0, c0,
                                              _ptr(src1, H16, "src1"), h1, w1, c1,
                                              _ptr(out, H16, "out"), B, Ho, Wo, _launch())
    _lib.check(rc, "spg_upsample_concat_h16", dn)


def fusion_combine(g2, g3, g4, bias, fused, row_sums, B: int, Hs: int, Cc: int) -> None:
    lib, dn = _lib_for(fused)
    rc = lib.spg_fusion_combine(_ptr(g2, torch.float32, "g2"), _ptr(g3, torch.float32, "g3"),
                                        _ptr(g4, torch.float32, "g4"), _ptr(bias, torch.float32, "bias"),
                                        _ptr(fused, H16, "fused"), _ptr(row_sums, torch.float32, "row_sums"),
                                        B, Hs, Cc, _launch())
    _lib.check(rc, "spg_fusion_combine", dn)


def row_sums(x, out, B: int, H: int, W: int, Cc: int) -> None:
    lib, dn = _lib_for(x)
    rc = lib.spg_row_sums_h16(_ptr(x, H16, "x"), _ptr(out, torch.float32, "row_sums"), B, H, W, Cc,
                                       _launch())
    _lib.check(rc, "spg_row_sums_h16", dn)


def pooled_mlp(row_sums_t, rows: int, count: int, w1, b1, R: int, w2, out, B: int, Cc: int) -> None:
    lib, dn = _default_lib(row_sums_t)
    rc = lib.spg_pooled_mlp(_ptr(row_sums_t, torch.float32, "row_sums"), rows, count,
                                    _ptr(w1, torch.float32, "w1"), _ptr(b1, torch.float32, "b1"), R,
                                    _ptr(w2, torch.float32, "w2"), _ptr(out, torch.float32, "out"), B, Cc, _launch())
    _lib.check(rc, "spg_pooled_mlp", dn)


def scale_channels(x, gate, B: int, HW: int, Cc: int) -> None:
    lib, dn = _lib_for(x)
    rc = lib.spg_scale_channels_h16(_ptr(x, H16, "x"), _ptr(gate, torch.float32, "gate"), B, HW, Cc,
                                             _launch())
    _lib.check(rc, "spg_scale_channels_h16", dn)


def easpp_branches(x, dw, dw_bias, gvec, wf, wf_bias, y, B: int, H: int, W: int, dilations) -> None:
    dil = (C.c_int * 4)(*dilations)
    lib, dn = _lib_for(x)
    rc = lib.spg_easpp_branches(_ptr(x, H16, "x"), _ptr(dw, torch.float32, "dw"),
                                        _ptr(dw_bias, torch.float32, "dw_bias"), _ptr(gvec, torch.float32, "gvec"),
                                        _ptr(wf, torch.float32, "wf"), _ptr(wf_bias, torch.float32, "wf_bias"),
                                        _ptr(y, H16, "y"), B, H, W, dil, _launch())
    _lib.check(rc, "spg_easpp_branches", dn)


def nhwc_to_nchw_f32(x: torch.Tensor, y: torch.Tensor, B: int, HW: int, Cc: int) -> None:
    lib, dn = _lib_for(x)
    rc = lib.spg_nhwc_h16_to_nchw_f32(_ptr(x, H16, "x"), _ptr(y, torch.float32, "y"), B, HW, Cc,
                                               _launch())
    _lib.check(rc, "spg_nhwc_h16_to_nchw_f32", dn)


def mask_stats(logits: torch.Tensor, gt_u8: torch.Tensor, double_sigmoid: bool = False):
    """logits [B,1,H,W] fp32, gt_u8 [B,H,W] uint8 (foreground > 128) -> (mask uint8 [B,H,W], stats int64 [B,5]).
    See spg_mask_stats_u8; `mae_from_stats` turns the statistics into the reference's MAE."""
    B = logits.shape[0]
    HW = logits[0].numel()
    mask = torch.empty(B, *gt_u8.shape[1:], dtype=torch.uint8, device=logits.device)
    stats = torch.empty(B, 8, dtype=torch.int32, device=logits.device)
    lib, dn = _default_lib(logits)
    rc = lib.spg_mask_stats_u8(_ptr(logits, torch.float32, "logits"), _ptr(gt_u8, torch.uint8, "gt"),
                               _ptr(mask, torch.uint8, "mask"), _ptr(stats, torch.int32, "stats"), B, HW,
                               int(double_sigmoid), _launch())
    _lib.check(rc, "spg_mask_stats_u8", dn)
    return mask, stats[:, :5].to(torch.int64)


def mae_from_stats(stats: torch.Tensor, n_pixels: int) -> torch.Tensor:
    """Per-image MAE of py_sod_metrics (pred / 255, min-max normalised unless constant; gt > 128) from the integer
    statistics of `mask_stats`, in fp64: exact, and identical on any number of GPUs."""
    s = stats.to(torch.float64)
    lo, hi, nfg, sbg, sfg = 255.0 - s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 4]
    nbg = n_pixels - nfg
    span = hi - lo
    norm = ((sbg - lo * nbg) + (hi * nfg - sfg)) / torch.where(span > 0, span, torch.ones_like(span))
    flat = sbg / 255.0 + (nfg - sfg / 255.0)
    return torch.where(span > 0, norm, flat) / n_pixels


def sod_gt_prepare(gt_u8: torch.Tensor):
    """gt_u8 [B,H,W] uint8 (foreground > 128) -> (nearest int32 [B,H,W], gt_stats int64 [B,4]).  Ground truth only:
    cache the result per dataset.  See spg_sod_gt_prepare_u8."""
    B, H, W = gt_u8.shape
    lib, dn = _default_lib(gt_u8)
    nbytes = int(lib.spg_sod_workspace_bytes(B, H, W))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=gt_u8.device)
    nearest = torch.empty(B, H, W, dtype=torch.int32, device=gt_u8.device)
    stats = torch.empty(B, 4, dtype=torch.int64, device=gt_u8.device)
    rc = lib.spg_sod_gt_prepare_u8(_ptr(gt_u8, torch.uint8, "gt"), B, H, W, _ptr(nearest, torch.int32, "nearest"),
                                   _ptr(stats, torch.int64, "gt_stats"), _ptr(ws, torch.uint8, "workspace"), nbytes,
                                   _launch())
    _lib.check(rc, "spg_sod_gt_prepare_u8", dn)
    return nearest, stats


def sod_scores(pred_u8: torch.Tensor, gt_u8: torch.Tensor, nearest: torch.Tensor, gt_stats: torch.Tensor) -> torch.Tensor:
    """uint8 masks + prepared ground truth -> fp64 [B,5] = (S-alpha, weighted F, MAE, adaptive E, mean F) per image
    (the per-sample scores of utils/metrics.py:161-167).  See spg_sod_scores_u8."""
    B, H, W = gt_u8.shape
    if tuple(pred_u8.shape) != (B, H, W):
        raise ValueError(f"pred {tuple(pred_u8.shape)} and gt {tuple(gt_u8.shape)} differ")
    lib, dn = _default_lib(pred_u8)
    nbytes = int(lib.spg_sod_workspace_bytes(B, H, W))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=gt_u8.device)
    scores = torch.empty(B, 5, dtype=torch.float64, device=gt_u8.device)
    rc = lib.spg_sod_scores_u8(_ptr(pred_u8, torch.uint8, "pred"), _ptr(gt_u8, torch.uint8, "gt"),
                               _ptr(nearest, torch.int32, "nearest"), _ptr(gt_stats, torch.int64, "gt_stats"), B, H, W,
                               _ptr(scores, torch.float64, "scores"), _ptr(ws, torch.uint8, "workspace"), nbytes,
                               _launch())
    _lib.check(rc, "spg_sod_scores_u8", dn)
    return scores


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # utils/image_processor.py:67-68
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_rgb(img_u8: torch.Tensor, size: int, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 HWC RGB image on the device -> fp32 [3,size,size], the reference's process_image
    (utils/image_processor.py:114-134: /255, antialiased bilinear resize, ImageNet normalisation).  See
    spg_preprocess_rgb_u8."""
    if img_u8.dim() != 3 or img_u8.shape[2] != 3:
        raise ValueError(f"expected a uint8 [H,W,3] image, got {tuple(img_u8.shape)}")
    H, W, _ = img_u8.shape
    lib, dn = _default_lib(img_u8)
    if out is None:
        out = torch.empty(3, size, size, dtype=torch.float32, device=img_u8.device)
    nbytes = int(lib.spg_preprocess_workspace_bytes(H, W, size))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=img_u8.device)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    rc = lib.spg_preprocess_rgb_u8(_ptr(img_u8, torch.uint8, "img"), H, W, _ptr(out, torch.float32, "out"), size, m, s,
                                   _ptr(ws, torch.uint8, "workspace"), nbytes, _launch())
    _lib.check(rc, "spg_preprocess_rgb_u8", dn)
    return out


def resize_bilinear(src: torch.Tensor, size, sigmoid: bool = False) -> torch.Tensor:
    """fp32 [B,1,H,W] / [B,H,W] -> the same rank at `size`, F.interpolate(mode='bilinear', align_corners=False)
    [+ sigmoid]: engine/predictor.py:350-365, engine/evaluator.py:539-554.  See spg_resize_bilinear_f32."""
    ho, wo = int(size[0]), int(size[1])
    lead = src.shape[:-2]
    hi, wi = src.shape[-2:]
    B = 1
    for d in lead:
        B *= int(d)
    dst = torch.empty(*lead, ho, wo, dtype=torch.float32, device=src.device)
    lib, dn = _default_lib(src)
    rc = lib.spg_resize_bilinear_f32(_ptr(src, torch.float32, "src"), B, hi, wi, _ptr(dst, torch.float32, "dst"), ho, wo,
                                     int(sigmoid), _launch())
    _lib.check(rc, "spg_resize_bilinear_f32", dn)
    return dst
