"""Runs ONE kernel of the library at its batch-64 shape a few times (for `ncu -k regex:<name> -s 2 -c 1` captures):
    python tools/kernel_probe.py {up2conv|upcat|aspp|layernorm|fusion|patchify|attn_tc|attn_global|sod}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402
from spegnet_b200.model import up2_phase_weights  # noqa: E402

what = sys.argv[1]
B, dev, h16 = 64, "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731


def run(fn, n=4):
    for _ in range(n):
        fn()
    torch.cuda.synchronize()


if what == "up2conv":
    x = rn(B, 256, 256, 128).to(h16)
    main, dl, dr = up2_phase_weights(rn(64, 128, 3, 3) / 34.0)
    corr = torch.zeros(2, B * 256, 256, device=dev)
    out = torch.empty(B, 512, 512, 64, device=dev, dtype=h16)
    wp, b4 = main.to(h16).contiguous(), rn(256)
    run(lambda: ops.conv3x3_up2(x, wp, corr, b4, out))
elif what == "upcat":
    a, e = rn(B, 128, 128, 256).to(h16), rn(B, 64, 64, 64).to(h16)
    out = torch.empty(B, 256, 256, 320, device=dev, dtype=h16)
    run(lambda: ops.upsample_concat(a, e, out))
elif what == "aspp":
    x = rn(B, 64, 64, 128).to(h16)
    y = torch.empty_like(x)
    dw, dwb, gv, wf, wfb = rn(4, 9, 128), rn(4, 128), rn(B, 128), rn(128, 5), rn(128)
    run(lambda: ops.easpp_branches(x, dw, dwb, gv, wf, wfb, y, B, 64, 64, (1, 6, 12, 18)))
elif what == "layernorm":
    x = rn(65536, 576)
    y = torch.empty(65536, 576, device=dev, dtype=h16)
    ga, be = rn(576), rn(576)
    run(lambda: ops.layernorm(x, ga, be, y, 1e-6))
elif what == "fusion":
    g2, g3, g4 = rn(B * 4096, 512), rn(B * 1024, 512), rn(B * 256, 512)
    fused = torch.empty(B, 64, 64, 512, device=dev, dtype=h16)
    rs = torch.empty(B * 64 * 512, device=dev)
    bias = rn(512)
    run(lambda: ops.fusion_combine(g2, g3, g4, bias, fused, rs, B, 64, 512))
elif what == "patchify":
    x = rn(B, 3, 512, 512)
    cols = torch.empty(B * 16384, 168, device=dev, dtype=h16)
    run(lambda: ops.patchify(x, cols))
elif what in ("attn_tc", "attn_global"):
    qkv = (rn(B * 1024, 1728) * 1.5).to(h16)
    out = torch.empty(B * 1024, 576, device=dev, dtype=h16)
    run(lambda: ops.window_attention(qkv, out, B, 32, 32, 576, 8, 16 if what == "attn_tc" else 0, False))
elif what == "sod":
    pred = torch.randint(0, 256, (B, 512, 512), dtype=torch.uint8, device=dev, generator=g)
    yy, xx = torch.meshgrid(torch.arange(512, device=dev), torch.arange(512, device=dev), indexing="ij")
    gt = (((yy - 256) ** 2 / 9000.0 + (xx - 200) ** 2 / 20000.0) < 1).to(torch.uint8).mul(255)[None].repeat(B, 1, 1).contiguous()
    nearest, stats = ops.sod_gt_prepare(gt)
    run(lambda: ops.sod_scores(pred, gt, nearest, stats))
else:
    raise SystemExit(__doc__)
print("ok", what)
