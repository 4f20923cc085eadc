"""Times / profiles the attention kernel alone on the stage-3 windowed shape (B=64): python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402

B, H, D, heads = 64, 32, 576, 8
qkv = (torch.randn(B * H * H, 3 * D, device="cuda") * 1.5).half()
out = torch.empty(B * H * H, D, device="cuda", dtype=torch.float16)
for ws, name in ((16, "windowed 16x16"), (0, "global 32x32")):
    for _ in range(3):
        ops.window_attention(qkv, out, B, H, H, D, heads, ws, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.window_attention(qkv, out, B, H, H, D, heads, ws, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nk = (ws or H) ** 2
    flops = 4.0 * B * H * H * nk * D
    print(f"{name}: {ms * 1e3:.1f} us  {flops / ms * 1e-9:.1f} TFLOP/s")
