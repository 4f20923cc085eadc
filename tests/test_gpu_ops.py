"""Per-kernel parity: every C-ABI entry point against the plain fp32 PyTorch op it replaces, on a B200.
Inputs are rounded to bf16 first where the kernel takes bf16, so the tolerance only has to absorb
the output rounding (bf16 outputs: ~2^-8 relative) and fp32 summation order."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# Every test runs against both library variants (libspegnet_b200_fp16.so / _bf16.so).  Tolerances at the call
# sites are written for bf16 (2^-8 relative output rounding); a 16-bit result of the fp16 build is held to an
# 8x tighter bound (2^-11 rounding), so an fp16-only regression cannot hide behind the bf16 bound.
H16 = torch.bfloat16
FP16_TIGHTEN = 8.0


@pytest.fixture(scope="module", params=["fp16", "bf16"])
def ops(request):
    global H16
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spegnet_b200 import _lib, ops as _ops

    lib = _lib.load(request.param)
    assert lib.spg_device_check() == 0, lib.spg_last_error()
    H16 = torch.float16 if request.param == "fp16" else torch.bfloat16
    return _ops


def _bf(t):
    return t.to(H16)


def _close(got, ref, atol, rtol):
    if got.dtype == torch.float16:
        atol, rtol = atol / FP16_TIGHTEN, rtol / FP16_TIGHTEN
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.4e} (ref max {ref.abs().max().item():.3f})"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (384, 144, 144), (512, 864, 288), (300, 192, 128), (1024, 576, 2304),
                                   (1024, 512, 1152)])
def test_linear_plain(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=H16)
    ops.linear(a, w, out, bias=bias)
    _close(out, a.float() @ w.float().t() + bias, 2e-2, 1e-2)


def test_linear_gelu_residual_f32(ops):
    M, N, K = 1024, 576, 576
    g = torch.Generator(device="cuda").manual_seed(3)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = F.gelu(a.float() @ w.float().t() + bias) + res
    out = res.clone()
    ops.linear(a, w, out, bias=bias, act=ops.ACT_GELU, residual=out)  # in place, as the trunk uses it
    _close(out, ref, 2e-4, 1e-4)


def test_linear_broadcast_residual_and_head(ops):
    M, N, K = 1024, 144, 160
    g = torch.Generator(device="cuda").manual_seed(4)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    pos = torch.randn(256, N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    ops.linear(a, w, out, residual=pos, res_rows=256)
    _close(out, a.float() @ w.float().t() + pos.repeat(4, 1), 2e-4, 1e-4)
    hw = torch.randn(N, device="cuda", generator=g)
    head = torch.empty(M, device="cuda")
    ops.linear(a, w, None, act=ops.ACT_RELU, head_w=hw, head_b=0.5, head_out=head)
    _close(head, F.relu(a.float() @ w.float().t()) @ hw + 0.5, 5e-4, 1e-4)


def test_linear_rejects_bad_shapes(ops):
    a = torch.zeros(128, 60, device="cuda", dtype=H16)
    w = torch.zeros(64, 60, device="cuda", dtype=H16)
    with pytest.raises(ValueError):
        ops.linear(a, w, torch.empty(128, 64, device="cuda", dtype=H16))
    with pytest.raises(ValueError):
        ops.linear(a.float(), w, torch.empty(128, 64, device="cuda", dtype=H16))


@pytest.mark.parametrize("B,H,Cin,Cout,head", [(2, 16, 64, 64, False), (1, 64, 256, 64, True), (1, 128, 320, 256, True),
                                               (1, 256, 128, 128, False), (2, 512, 64, 64, True)])  # last: resident weights
def test_conv3x3(ops, B, H, Cin, Cout, head):
    g = torch.Generator(device="cuda").manual_seed(H + Cin)
    x = _bf(torch.randn(B, H, H, Cin, device="cuda", generator=g))
    w4 = _bf(torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cin))
    bias = torch.randn(Cout, device="cuda", generator=g)
    wk = w4.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = torch.empty(B * H * H, Cout, device="cuda", dtype=H16)
    hw = torch.randn(Cout, device="cuda", generator=g) if head else None
    ho = torch.empty(B, 1, H, H, device="cuda") if head else None
    ops.conv3x3(x, wk, out, bias=bias, act=ops.ACT_RELU, head_w=hw, head_b=-0.25, head_out=ho)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias, padding=1))
    _close(out.view(B, H, H, Cout).permute(0, 3, 1, 2), ref, 2e-2, 1e-2)
    if head:
        _close(ho, (ref * hw.view(1, -1, 1, 1)).sum(1, keepdim=True) - 0.25, 5e-3, 1e-3)


@pytest.mark.parametrize("C", [144, 288, 576, 1152])
def test_layernorm(ops, C):
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(1000, C, device="cuda", generator=g) * 3 + 1
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    y = torch.empty(1000, C, device="cuda", dtype=H16)
    ops.layernorm(x, gamma, beta, y, 1e-6)
    _close(y, F.layer_norm(x, (C,), gamma, beta, 1e-6), 2e-2, 1e-2)


def test_patch_embed(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    B, S = 2, 64
    x = torch.randn(B, 3, S, S, device="cuda", generator=g)
    w = torch.randn(144, 3, 7, 7, device="cuda", generator=g) / math.sqrt(147)
    bias = torch.randn(144, device="cuda", generator=g)
    cols = torch.empty(B * 256, 168, device="cuda", dtype=H16)
    ops.patchify(x, cols)
    # column order of spg_patchify_7x7s4: k = (ky*3 + c)*8 + kx, slot kx = 7 zero
    wk = F.pad(w.permute(0, 2, 1, 3), (0, 1)).reshape(144, 168).to(H16).contiguous()
    out = torch.empty(B * 256, 144, device="cuda")
    ops.linear(cols, wk, out, bias=bias)
    ref = F.conv2d(_bf(x).float(), _bf(w).float(), bias, stride=4, padding=3).permute(0, 2, 3, 1).reshape(-1, 144)
    _close(out, ref, 1e-3, 1e-3)
    assert float(cols.view(-1, 21, 8)[:, :, 7].abs().max()) == 0.0


def test_maxpool_and_cast(ops):
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(2, 16, 16, 288, device="cuda", generator=g)
    y = torch.empty(2, 8, 8, 288, device="cuda")
    ops.maxpool2x2(x, y, 2, 16, 16, 288)
    assert torch.equal(y, F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
    z = torch.empty(2 * 16 * 16 * 288, device="cuda", dtype=H16)
    ops.cast_h16(x.view(-1), z)
    assert torch.equal(z, x.view(-1).to(H16))


def _ref_attention(qkv, B, H, D, heads, ws, qpool):
    hd = D // heads
    t = qkv.float().view(B, H, H, 3 * D)
    ws = ws or H
    t = t.view(B, H // ws, ws, H // ws, ws, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, 3, heads, hd)
    q, k, v = t[:, :, :, 0], t[:, :, :, 1], t[:, :, :, 2]
    wq = ws
    if qpool:
        q = F.max_pool2d(q.reshape(-1, ws, ws, D).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        wq = ws // 2
        q = q.reshape(-1, wq, wq, heads, hd)
    nw = q.shape[0]
    q = q.reshape(nw, wq * wq, heads, hd).transpose(1, 2)
    k = k.reshape(nw, ws * ws, heads, hd).transpose(1, 2)
    v = v.reshape(nw, ws * ws, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(nw, wq, wq, D)
    Ho = H // 2 if qpool else H
    o = o.view(B, H // ws, H // ws, wq, wq, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Ho, Ho, D)
    return o


@pytest.mark.parametrize("H,heads,ws,qpool", [
    (32, 2, 8, False),    # stage 1
    (32, 4, 8, True),     # block 2: 16 pooled queries x 64 keys
    (16, 4, 4, False),    # stage 2
    (16, 8, 4, True),     # block 8: 4 pooled queries x 16 keys (CUDA-core path)
    (32, 8, 16, False),   # stage 3 windows
    (32, 8, 0, False),    # stage 3 global, 1024 keys: two-pass tcgen05 kernel, 8 key blocks of 128
    (64, 8, 0, False),    # stage 3 global at 1024x1024 input: 4096 keys, 2 grid rows per 128-token tile
    (32, 16, 16, True),   # block 44
    (16, 16, 8, False),   # stage 4
])
def test_window_attention(ops, H, heads, ws, qpool):
    B, D = 2, heads * 72
    g = torch.Generator(device="cuda").manual_seed(H * 100 + heads + ws)
    qkv = _bf(torch.randn(B * H * H, 3 * D, device="cuda", generator=g) * 1.5)
    Ho = H // 2 if qpool else H
    out = torch.full((B * Ho * Ho, D), float("nan"), device="cuda", dtype=H16)
    ops.window_attention(qkv, out, B, H, H, D, heads, ws, qpool)
    ref = _ref_attention(qkv, B, H, D, heads, ws, qpool)
    _close(out.view(B, Ho, Ho, D), ref, 2e-2, 2e-2)


def test_upsample_concat(ops):
    g = torch.Generator(device="cuda").manual_seed(8)
    a = _bf(torch.randn(2, 16, 16, 256, device="cuda", generator=g))
    e = _bf(torch.randn(2, 8, 8, 64, device="cuda", generator=g))
    out = torch.empty(2, 32, 32, 320, device="cuda", dtype=H16)
    ops.upsample_concat(a, e, out)
    up = lambda t: F.interpolate(t.float().permute(0, 3, 1, 2), size=(32, 32), mode="bilinear", align_corners=False)  # noqa: E731
    ref = torch.cat([up(a), up(e)], 1).permute(0, 2, 3, 1)
    _close(out, ref, 1e-2, 1e-2)
    out2 = torch.empty(2, 32, 32, 256, device="cuda", dtype=H16)
    ops.upsample_concat(a, None, out2)
    _close(out2, up(a).permute(0, 2, 3, 1), 1e-2, 1e-2)


@pytest.mark.parametrize("h0,w0,c0,h1,w1,c1,Ho,Wo", [
    (16, 16, 256, 16, 16, 64, 32, 32),   # decoder stage 1: both sources x2 (constant-tap kernel)
    (16, 24, 128, 8, 12, 64, 32, 48),    # stage 2 ratios (x2 + x4) on a non-square grid
    (8, 8, 64, 8, 8, 8, 16, 16),         # smallest grid: every pixel is a border pixel of the 3x3 / 2x2 patches
    (8, 8, 64, 16, 16, 32, 32, 32),      # src0 x4: generic kernel
    (4, 6, 32, 0, 0, 0, 12, 18),         # x3, no second source: generic kernel
])
def test_upsample_concat_ratios(ops, h0, w0, c0, h1, w1, c1, Ho, Wo):
    """Every pixel (borders included) of both kernels against ATen's bilinear on the same 16-bit inputs: the result is
    the 16-bit rounding of an fp32 interpolation, so it may differ from the rounded reference by one ulp at most."""
    if Ho % 2 or Wo % 2:
        pytest.skip("even outputs only")
    g = torch.Generator(device="cuda").manual_seed(h0 * 7 + Wo)
    a = _bf(torch.randn(2, h0, w0, c0, device="cuda", generator=g))
    e = _bf(torch.randn(2, h1, w1, c1, device="cuda", generator=g)) if c1 else None
    out = torch.full((2, Ho, Wo, c0 + c1), float("nan"), device="cuda", dtype=H16)
    ops.upsample_concat(a, e, out)
    up = lambda t: F.interpolate(t.float().permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=False)  # noqa: E731
    ref = torch.cat([up(a)] + ([up(e)] if c1 else []), 1).permute(0, 2, 3, 1)
    ulp = 2.0 ** -10 if H16 == torch.float16 else 2.0 ** -7
    err = (out.float() - ref).abs()
    assert bool((err <= ulp * ref.abs().clamp_min(2.0 ** -14) * 1.01 + 1e-7).all()), f"max err {err.max().item():.3e}"


@pytest.mark.parametrize("Hs", [16, 8, 4, 64])
def test_fusion_combine_se_scale(ops, Hs):
    g = torch.Generator(device="cuda").manual_seed(9 + Hs)
    B, C = 2, 512
    g2 = torch.randn(B, Hs, Hs, C, device="cuda", generator=g)
    g3 = torch.randn(B, Hs // 2, Hs // 2, C, device="cuda", generator=g)
    g4 = torch.randn(B, Hs // 4, Hs // 4, C, device="cuda", generator=g)
    bias = torch.randn(C, device="cuda", generator=g)
    fused = torch.empty(B, Hs, Hs, C, device="cuda", dtype=H16)
    rs = torch.empty(B, Hs, C, device="cuda")
    ops.fusion_combine(g2, g3, g4, bias, fused, rs, B, Hs, C)
    up = lambda t: F.interpolate(t.permute(0, 3, 1, 2), size=(Hs, Hs), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)  # noqa: E731
    ref = F.relu(g2 + up(g3) + up(g4) + bias)
    _close(fused, ref, 1e-2, 1e-2)
    _close(rs, ref.sum(2), 1e-3, 1e-4)
    w1 = torch.randn(32, C, device="cuda", generator=g) / math.sqrt(C)
    w2 = torch.randn(C, 32, device="cuda", generator=g) / math.sqrt(32)
    gate = torch.empty(B, C, device="cuda")
    ops.pooled_mlp(rs, Hs, Hs * Hs, w1, None, 32, w2, gate, B, C)
    ref_gate = torch.sigmoid(F.relu(ref.mean((1, 2)) @ w1.t()) @ w2.t())
    _close(gate, ref_gate, 1e-5, 1e-4)
    before = fused.clone()
    ops.scale_channels(fused, gate, B, Hs * Hs, C)
    _close(fused, before.float() * gate[:, None, None, :], 1e-2, 1e-2)


def test_easpp(ops):
    g = torch.Generator(device="cuda").manual_seed(10)
    B, H = 2, 64
    x = _bf(F.relu(torch.randn(B, H, H, 128, device="cuda", generator=g)))
    rs = torch.empty(B, H, 128, device="cuda")
    ops.row_sums(x, rs, B, H, H, 128)
    _close(rs, x.float().sum(2), 1e-3, 1e-4)
    wg = torch.randn(128, 128, device="cuda", generator=g) / math.sqrt(128)
    bg = torch.randn(128, device="cuda", generator=g)
    gvec = torch.empty(B, 128, device="cuda")
    ops.pooled_mlp(rs, H, H * H, wg, bg, 128, None, gvec, B, 128)
    ref_g = F.relu(x.float().mean((1, 2)) @ wg.t() + bg)
    _close(gvec, ref_g, 1e-5, 1e-4)
    dil = (1, 6, 12, 18)
    dw = torch.randn(4, 128, 3, 3, device="cuda", generator=g) / 3
    dwb = torch.randn(4, 128, device="cuda", generator=g)
    wf = torch.randn(128, 5, device="cuda", generator=g) / math.sqrt(5)
    wfb = torch.randn(128, device="cuda", generator=g)
    y = torch.empty(B, H, H, 128, device="cuda", dtype=H16)
    dw_k = dw.reshape(4, 128, 9).permute(0, 2, 1).contiguous()  # [4, tap, ch]
    ops.easpp_branches(x, dw_k, dwb, ref_g.contiguous(), wf, wfb, y, B, H, H, dil)
    xn = x.float().permute(0, 3, 1, 2)
    br = [F.relu(F.conv2d(xn, dw[i].unsqueeze(1), dwb[i], padding=d, dilation=d, groups=128)) for i, d in enumerate(dil)]
    br.append(ref_g[:, :, None, None].expand(-1, -1, H, H))
    ref = F.relu(F.conv2d(torch.cat(br, 1), wf.view(128, 5, 1, 1), wfb, groups=128)).permute(0, 2, 3, 1)
    _close(y, ref, 2e-2, 1e-2)


def test_nhwc_to_nchw(ops):
    x = _bf(torch.randn(2, 8, 8, 64, device="cuda"))
    y = torch.empty(2, 64, 8, 8, device="cuda")
    ops.nhwc_to_nchw_f32(x, y, 2, 64, 64)
    assert torch.equal(y, x.float().permute(0, 3, 1, 2))


@pytest.mark.parametrize("double_sigmoid", [False, True])
def test_mask_stats_match_reference_quantisation_and_mae(ops, double_sigmoid):
    """uint8 mask + integer statistics on the GPU vs the oracle's quantisation and MAE (utils/metrics.py:205-210)."""
    import numpy as np

    from oracle import sod_metrics as M

    g = torch.Generator().manual_seed(12)
    logits = torch.randn(3, 1, 64, 64, generator=g) * 2.5 - 1.0
    gt = (torch.rand(3, 64, 64, generator=g) > 0.7).to(torch.uint8) * 255
    logits[2] = 0.3  # a constant prediction exercises the "no min-max normalisation" branch
    mask, stats = ops.mask_stats(logits.cuda(), gt.cuda(), double_sigmoid)
    mae = ops.mae_from_stats(stats, 64 * 64).cpu().numpy()
    for i in range(3):
        x = logits[i, 0].numpy()
        if double_sigmoid:
            x = 1 / (1 + np.exp(-x))
        ref_q = M.quantise_like_reference(x)
        got_q = mask[i].cpu().numpy()
        # expf vs numpy exp may disagree by one grey level exactly at a truncation boundary
        assert np.abs(got_q.astype(int) - ref_q.astype(int)).max() <= 1
        assert (got_q != ref_q).mean() < 2e-3
        ref_mae = M.score_pair(got_q, gt[i].numpy())["mae"]  # same mask -> the statistics must give the exact MAE
        assert abs(mae[i] - ref_mae) < 1e-12


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 8, 128, 64, 64), (1, 16, 256, 128, 64), (3, 5, 128, 64, 32),
                                             (8, 64, 256, 128, 64), (1, 12, 176, 128, 64), (2, 6, 96, 64, 64),
                                             (1, 9, 192, 128, 64)])  # last three: ragged low-resolution rows
def test_conv3x3_up2_matches_interpolate_then_conv(ops, B, H, W, Cin, Cout):
    """Fused bilinear x2 + 3x3 conv (+ folded bias, ReLU) against F.interpolate -> F.conv2d in fp32
    (models/object_detection.py:219,230-232); the last shape is large enough for the CTA-pair instance."""
    from spegnet_b200.model import up2_phase_weights

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    x = _bf(torch.randn(B, H, W, Cin, device="cuda", generator=g))
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cin)
    bias = torch.randn(Cout, device="cuda", generator=g)
    main, dwl, dwr = up2_phase_weights(w)
    bord = torch.empty(2, B * H, 9 * Cin, device="cuda", dtype=H16)
    corr = torch.empty(2, B * H, 4 * Cout, device="cuda", dtype=torch.float32)
    out = torch.empty(B, 2 * H, 2 * W, Cout, device="cuda", dtype=H16)
    ops.up2_border_gather(x, bord)
    ops.linear(bord[0], _bf(dwl).contiguous(), corr[0])
    ops.linear(bord[1], _bf(dwr).contiguous(), corr[1])
    ops.conv3x3_up2(x, _bf(main).contiguous(), corr, bias.repeat(4).contiguous(), out)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False)
    ref = F.relu(F.conv2d(up, w, bias, padding=1)).permute(0, 2, 3, 1)
    _close(out, ref, 2e-2, 1e-2)
    # the border rows / columns are where the folded weights differ: check them on their own
    for sl in (out[:, :2], out[:, -2:], out[:, :, :2], out[:, :, -2:]):
        assert torch.isfinite(sl.float()).all()
    _close(out[:, :, :2], ref[:, :, :2], 2e-2, 1e-2)
    _close(out[:, :, -2:], ref[:, :, -2:], 2e-2, 1e-2)


@pytest.mark.parametrize("M,C,N2", [(1024, 576, 1728), (4096, 144, 576), (300, 288, 864), (65536, 576, 2304)])
def test_layernorm_folded_into_producer_and_consumer(ops, M, C, N2):
    """LayerNorm folded into the GEMMs around it: a residual GEMM (producer) emits the centred 16-bit copy + row
    records, a second GEMM (consumer) applies rstd / mean in its epilogue.  Reference: the unfused chain
    x = res + a W1^T + b1;  y = LayerNorm(x) * gamma + beta;  out = gelu(y W2^T + b2)  in fp32."""
    g = torch.Generator(device="cuda").manual_seed(M + C)
    K1 = 128
    a = _bf(torch.randn(M, K1, device="cuda", generator=g))
    w1 = _bf(torch.randn(C, K1, device="cuda", generator=g) / math.sqrt(K1))
    b1 = torch.randn(C, device="cuda", generator=g)
    # residual rows with a large common offset: the centring by the previous mean must absorb it
    res = torch.randn(M, C, device="cuda", generator=g) + 6.0 * torch.randn(M, 1, device="cuda", generator=g)
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.3
    w2 = torch.randn(N2, C, device="cuda", generator=g) / math.sqrt(C)
    b2 = torch.randn(N2, device="cuda", generator=g)
    # records of the residual input rows, as an earlier producer would have left them: {c = 0, P = 1, (sum, sum sq)}
    prev = torch.zeros(M, 32, device="cuda")
    prev[:, 1] = 1.0
    prev[:, 2] = res.sum(1)
    prev[:, 3] = (res * res).sum(1)
    x = torch.empty(M, C, device="cuda")
    x.copy_(res)
    xc = torch.empty(M, C, device="cuda", dtype=H16)
    rec = torch.full((M, 32), float("nan"), device="cuda")
    ops.linear(a, w1, x, bias=b1, residual=x, ln_emit=(rec, prev, xc))
    x_ref = res + a.float() @ w1.float().t() + b1
    _close(x, x_ref, 2e-2, 1e-2)
    centre = res.mean(1, keepdim=True)
    _close(xc, x_ref - centre, 2e-2, 1e-2)
    assert torch.allclose(rec[:, 0], centre[:, 0], atol=1e-4)
    parts = rec[:, 1].long()
    assert int(parts.min()) == int(parts.max()) and 2 <= int(parts[0]) <= 15
    P = int(parts[0])
    s1 = rec[:, 2:2 + 2 * P:2].sum(1)
    s2 = rec[:, 3:3 + 2 * P:2].sum(1)
    d = x_ref - centre
    assert torch.allclose(s1, d.sum(1), atol=2e-2, rtol=1e-3) and torch.allclose(s2, (d * d).sum(1), atol=5e-2, rtol=2e-3)
    # consumer: W' = W diag(gamma) (16-bit), colsum(W') of the ROUNDED W', bias' = b + W beta
    w2p = _bf(w2 * gamma[None, :]).contiguous()
    cw = w2p.float().sum(1).contiguous()
    b2p = (b2 + w2 @ beta).contiguous()
    for act, out_dtype in ((ops.ACT_GELU, H16), (ops.ACT_NONE, H16), (ops.ACT_NONE, torch.float32)):
        out = torch.empty(M, N2, device="cuda", dtype=out_dtype)
        ops.linear(xc, w2p, out, bias=b2p, act=act, ln_fold=(rec, cw, C, 1e-6))
        y = F.layer_norm(x_ref, (C,), gamma, beta, 1e-6)
        ref = y @ w2.t() + b2
        if act == ops.ACT_GELU:
            ref = F.gelu(ref)
        _close(out, ref, 3e-2, 1.5e-2)


@pytest.mark.parametrize("B,H,heads,ws", [(2, 32, 8, 16), (2, 32, 8, 0), (1, 64, 8, 0), (20, 32, 8, 0)])
def test_tcgen05_attention_kernels_directly(ops, B, H, heads, ws):
    """The tcgen05 / TMEM kernels through their own entry point: 16x16 windows, global 1024 keys, global 4096 keys (2 grid rows per 128-token tile), and a
    batch with more items than SMs (persistent loop over items, ring wrap-around)."""
    D = heads * 72
    g = torch.Generator(device="cuda").manual_seed(B * 7 + H + ws)
    qkv = _bf(torch.randn(B * H * H, 3 * D, device="cuda", generator=g) * 1.5)
    out = torch.full((B * H * H, D), float("nan"), device="cuda", dtype=H16)
    ops.window_attention_tc(qkv, out, B, H, H, D, heads, ws, False)
    ref = _ref_attention(qkv, B, H, D, heads, ws, False)
    _close(out.view(B, H, H, D), ref, 2e-2, 2e-2)


def test_fp16_stores_saturate(ops):
    """Every 16-bit store converts with .satfinite (csrc/half16.cuh): an accumulator beyond the fp16 range becomes
    +-65504, never inf (which the next LayerNorm / softmax would turn into NaN).  bf16 has fp32's range: finite too."""
    M, N, K = 128, 64, 64
    a = torch.full((M, K), 1000.0, device="cuda").to(H16)
    w = torch.full((N, K), 100.0, device="cuda").to(H16)
    w[N // 2:] *= -1
    out = torch.empty(M, N, device="cuda", dtype=H16)
    ops.linear(a, w, out)
    assert bool(torch.isfinite(out.float()).all())
    if H16 == torch.float16:
        assert float(out.float().max()) == 65504.0 and float(out.float().min()) == -65504.0
    else:
        _close(out, a.float() @ w.float().t(), 0.0, 1e-2)
    y = torch.empty(8, 144, device="cuda", dtype=H16)
    x = torch.zeros(8, 144, device="cuda")
    x[:, 0] = 1.0
    ops.layernorm(x, torch.full((144,), 1e4, device="cuda"), torch.zeros(144, device="cuda"), y, 1e-6)  # 1e4 * 11.96
    assert bool(torch.isfinite(y.float()).all())
    z = torch.empty(1024, device="cuda", dtype=H16)
    ops.cast_h16(torch.full((1024,), 1e6, device="cuda"), z)
    assert bool(torch.isfinite(z.float()).all())


@pytest.mark.parametrize("M,N,K,res_rows", [(16384, 576, 2304, 0), (12928, 576, 2304, 0), (16384, 576, 576, 0), (1000, 576, 576, 0),
                                            (4096, 288, 1152, 0), (20000, 288, 288, 0), (2000, 144, 576, 0),
                                            (4096, 144, 168, 1024), (128, 576, 2304, 0)])
def test_layernorm_applied_by_the_producer(ops, M, N, K, res_rows):
    """spg_epilogue_t.ln_apply_*: the residual GEMM stores the fp32 stream AND y = LayerNorm(stream) * gamma + beta as
    16 bit; the CTAs (pairs for long K) holding the n-tiles of a row block exchange row statistics through distributed
    shared memory.  Covers clusters of 1 / 2 / 3 CTAs and of 3 CTA pairs, a ragged last row block and the broadcast
    residual of the patch embedding."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(res_rows or M, N, device="cuda", generator=g) * 2 + 0.7
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    beta = torch.randn(N, device="cuda", generator=g)
    ref_x = a.float() @ w.float().t() + bias + (res.repeat(M // res_rows, 1) if res_rows else res)
    out = torch.empty(M, N, device="cuda") if res_rows else res.clone()
    y = torch.full((M, N), float("nan"), device="cuda", dtype=H16)
    ops.linear(a, w, out, bias=bias, residual=res if res_rows else out, res_rows=res_rows, ln_apply=(gamma, beta, y, 1e-6))
    _close(out, ref_x, 3e-4, 1e-4)
    _close(y, F.layer_norm(ref_x, (N,), gamma, beta, 1e-6), 2e-2, 1e-2)
    # the statistics are combined in a fixed order: a second launch and a launch over a row subset give the same bits
    out2 = torch.empty(M, N, device="cuda") if res_rows else res.clone()
    y2 = torch.empty_like(y)
    ops.linear(a, w, out2, bias=bias, residual=res if res_rows else out2, res_rows=res_rows, ln_apply=(gamma, beta, y2, 1e-6))
    assert torch.equal(y, y2)
    if M >= 256 and not res_rows:
        out3 = res[:128].clone()
        y3 = torch.empty(128, N, device="cuda", dtype=H16)
        ops.linear(a[:128].contiguous(), w, out3, bias=bias, residual=out3, ln_apply=(gamma, beta, y3, 1e-6))
        assert torch.equal(y3, y[:128])


@pytest.mark.parametrize("C", [144, 288, 576])
def test_matched_layernorm_is_bit_identical_to_the_producer(ops, C):
    """spg_layernorm_matched_f32_h16 on the fp32 stream == the y a residual GEMM stores with ln_apply, bit for bit (the
    host switches between the two by batch size)."""
    M, K = 1000, 4 * C
    g = torch.Generator(device="cuda").manual_seed(C)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(C, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(C, device="cuda", generator=g)
    out = torch.randn(M, C, device="cuda", generator=g) * 3 + 0.3
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    y = torch.empty(M, C, device="cuda", dtype=H16)
    ops.linear(a, w, out, bias=bias, residual=out, ln_apply=(gamma, beta, y, 1e-6))
    y2 = torch.empty_like(y)
    ops.layernorm_matched(out, gamma, beta, y2, 1e-6)
    assert torch.equal(y, y2)
    _close(y2, F.layer_norm(out, (C,), gamma, beta, 1e-6), 2e-2, 1e-2)


class _Guarded:
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_unavailable.txt), so out-of-bounds WRITES are
    caught the manual way: every output lives between two guard bands filled with a sentinel bit pattern."""

    GUARD = 4096  # elements on each side

    def __init__(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        self.flat = torch.empty(n + 2 * self.GUARD, dtype=dtype, device="cuda")
        self.sentinel = -7.0 if dtype.is_floating_point else 171
        self.flat.fill_(self.sentinel)
        self.view = self.flat[self.GUARD:self.GUARD + n].view(*shape)

    def intact(self):
        lo, hi = self.flat[:self.GUARD], self.flat[-self.GUARD:]
        return bool((lo == self.sentinel).all()) and bool((hi == self.sentinel).all())


def test_no_writes_outside_the_outputs(ops):
    g = torch.Generator(device="cuda").manual_seed(99)
    # GEMM with ragged M / N tails, fp32 residual, 16-bit and fp32 outputs, producer-side LayerNorm (16-bit second output)
    for M, N, K in [(300, 192, 128), (1000, 576, 576), (129, 144, 168), (2000, 1728, 576)]:
        a = _bf(torch.randn(M, K, device="cuda", generator=g))
        w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
        out = _Guarded((M, N), H16)
        ops.linear(a, w, out.view, bias=torch.randn(N, device="cuda", generator=g), act=ops.ACT_GELU)
        assert out.intact(), ("linear h16", M, N, K)
        outf = _Guarded((M, N), torch.float32)
        outf.view.normal_(generator=g)
        if N in (144, 576):
            y = _Guarded((M, N), H16)
            ops.linear(a, w, outf.view, residual=outf.view, ln_apply=(torch.ones(N, device="cuda"), torch.zeros(N, device="cuda"), y.view, 1e-6))
            assert y.intact() and bool(torch.isfinite(y.view.float()).all()), ("ln_apply", M, N, K)
            y2 = _Guarded((M, N), H16)
            ops.layernorm_matched(outf.view, torch.ones(N, device="cuda"), torch.zeros(N, device="cuda"), y2.view, 1e-6)
            assert y2.intact() and torch.equal(y2.view, y.view)
        else:
            ops.linear(a, w, outf.view, residual=outf.view)
        assert outf.intact(), ("linear f32", M, N, K)
    # implicit-GEMM conv with fused head, fused up2 conv
    x = _bf(torch.randn(1, 64, 64, 64, device="cuda", generator=g))
    wk = _bf(torch.randn(64, 9 * 64, device="cuda", generator=g) / 24)
    out = _Guarded((64 * 64, 64), H16)
    head = _Guarded((1, 1, 64, 64), torch.float32)
    ops.conv3x3(x, wk, out.view, bias=torch.zeros(64, device="cuda"), act=ops.ACT_RELU, head_w=torch.ones(64, device="cuda"), head_out=head.view)
    assert out.intact() and head.intact()
    # LayerNorm, attention (windowed tcgen05, small windows, pooled queries), max pool, cast
    xs = torch.randn(1000, 576, device="cuda", generator=g)
    y = _Guarded((1000, 576), H16)
    ops.layernorm(xs, torch.ones(576, device="cuda"), torch.zeros(576, device="cuda"), y.view, 1e-6)
    assert y.intact()
    for (B, Hh, D, heads, win, pool) in [(1, 32, 576, 8, 16, False), (1, 32, 144, 2, 8, False), (2, 16, 288, 4, 8, True), (1, 16, 288, 4, 4, False),
                                         (1, 32, 576, 8, 0, False)]:
        qkv = _bf(torch.randn(B * Hh * Hh, 3 * D, device="cuda", generator=g))
        Ho = Hh // 2 if pool else Hh
        o = _Guarded((B * Ho * Ho, D), H16)
        ops.window_attention(qkv, o.view, B, Hh, Hh, D, heads, win, pool)
        assert o.intact() and bool(torch.isfinite(o.view.float()).all()), (B, Hh, D, heads, win, pool)
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,act", [(40000, 432, 144, 0), (30001, 576, 144, 2), (80000, 144, 168, 0), (33000, 864, 144, 0),
                                       (26000, 288, 144, 0), (20000, 864, 288, 0), (45000, 288, 288, 0)])
def test_linear_resident_weights(ops, M, N, K, act):
    """Short-K GEMMs with many tiles run in the resident-weight mode (the CTA's weight tile is loaded once, the ring
    holds A chunks only, grid = a multiple of the n-tiles): ragged M, K tails (144 = 2*64 + 16), GELU, fp32 output."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    if act == 2:
        ref = F.gelu(ref)
    out = torch.empty(M, N, device="cuda", dtype=H16)
    ops.linear(a, w, out, bias=bias, act=act)
    _close(out, ref, 2e-2, 1e-2)
    outf = torch.empty(M, N, device="cuda")
    ops.linear(a, w, outf, bias=bias, act=act)
    _close(outf, ref, 3e-4, 1e-4)


@pytest.mark.parametrize("B,H,W,Cin,Cout,head", [(2, 44, 44, 256, 64, True), (1, 88, 88, 320, 256, True), (1, 176, 176, 128, 128, False),
                                                 (1, 48, 48, 64, 64, False), (2, 96, 96, 64, 128, True)])
def test_conv3x3_ragged_widths(ops, B, H, W, Cin, Cout, head):
    """Widths that neither divide 128 nor are multiples of it (the head at 352 / 384 inputs): 64- or 128-pixel tile
    columns whose last one is ragged -- A rows beyond W are TMA zero fill, the NHWC store map clips them, the fused head
    writes the compact map."""
    g = torch.Generator(device="cuda").manual_seed(H + Cin)
    x = _bf(torch.randn(B, H, W, Cin, device="cuda", generator=g))
    w4 = _bf(torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cin))
    bias = torch.randn(Cout, device="cuda", generator=g)
    wk = w4.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = _Guarded((B * H * W, Cout), H16)
    hw = torch.randn(Cout, device="cuda", generator=g) if head else None
    ho = _Guarded((B, 1, H, W), torch.float32) if head else None
    ops.conv3x3(x, wk, out.view, bias=bias, act=ops.ACT_RELU, head_w=hw, head_b=-0.25, head_out=ho.view if head else None)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias, padding=1))
    assert out.intact()
    _close(out.view.view(B, H, W, Cout).permute(0, 3, 1, 2), ref, 2e-2, 1e-2)
    if head:
        assert ho.intact()
        _close(ho.view, (ref * hw.view(1, -1, 1, 1)).sum(1, keepdim=True) - 0.25, 5e-3, 1e-3)
        ho2 = torch.empty(B, 1, H, W, device="cuda")
        ops.conv3x3(x, wk, None, bias=bias, act=ops.ACT_RELU, head_w=hw, head_b=-0.25, head_out=ho2)  # head only
        assert torch.equal(ho2, ho.view)


def test_copy_grid_pads_and_crops(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    src = _bf(torch.randn(2, 22, 22, 576, device="cuda", generator=g))
    dst = torch.zeros(2, 32, 32, 576, device="cuda", dtype=H16)
    ops.copy_grid(src, dst, 22, 22)
    assert torch.equal(dst[:, :22, :22], src) and float(dst[:, 22:].abs().max()) == 0 and float(dst[:, :, 22:].abs().max()) == 0
    back = torch.empty_like(src)
    ops.copy_grid(dst, back, 22, 22)
    assert torch.equal(back, src)
