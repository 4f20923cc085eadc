"""Times the three upsample+concat launches of the decoder at B=64: python tools/upcat_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402

B = 64
for (h, c0, he, c1, ho) in ((64, 256, 64, 64, 128), (128, 256, 64, 64, 256), (256, 128, 0, 0, 512)):
    a = torch.randn(B, h, h, c0, device="cuda").half()
    e = torch.randn(B, he, he, c1, device="cuda").half() if c1 else None
    out = torch.empty(B, ho, ho, c0 + c1, device="cuda", dtype=torch.float16)
    for _ in range(2):
        ops.upsample_concat(a, e, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.upsample_concat(a, e, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (out.numel() + a.numel() + (e.numel() if e is not None else 0)) * 2 / 1e9
    print(f"{h}->{ho} C={c0 + c1}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
