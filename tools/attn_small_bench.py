"""Times the small-window attention launches of stages 1-2 at B=64 (python tools/attn_small_bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402

B = 64
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for (H, D, heads, ws, pool) in ((128, 144, 2, 8, False), (128, 288, 4, 8, True), (64, 288, 4, 4, False), (64, 576, 8, 4, True)):
    Din = D
    qkv = (torch.randn(B * H * H, 3 * D, device="cuda") * 1.0).half()
    Ho = H // 2 if pool else H
    out = torch.empty(B * Ho * Ho, D, device="cuda", dtype=torch.float16)
    for _ in range(2):
        ops.window_attention(qkv, out, B, H, H, D, heads, ws, pool)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.window_attention(qkv, out, B, H, H, D, heads, ws, pool)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 5
    nbytes = (qkv.numel() + out.numel()) * 2
    print(f"H={H} D={D} heads={heads} ws={ws} pool={pool}: {ms * 1e3:.1f} us  {nbytes / ms * 1e-6:.0f} GB/s ({nbytes / 1e6:.0f} MB)", flush=True)
