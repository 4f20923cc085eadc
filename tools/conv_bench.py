"""Stand-alone timing of the decoder convolutions at batch 64 (development aid; GPU box).  [SPG_CONV_BRES=0] python tools/conv_bench.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops

H16 = torch.float16
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for (B, H, Cin, Cout, head, store) in [(64, 512, 64, 64, True, False), (64, 64, 256, 64, True, True), (64, 256, 128, 128, True, True),
                                       (64, 256, 320, 128, False, True), (64, 128, 320, 256, False, True)]:
    x = torch.randn(B, H, H, Cin, device="cuda").to(H16)
    w = (torch.randn(Cout, 9 * Cin, device="cuda") / (9 * Cin) ** 0.5).to(H16)
    bias = torch.randn(Cout, device="cuda")
    out = torch.empty(B * H * H, Cout, device="cuda", dtype=H16) if store else None
    hw = torch.randn(Cout, device="cuda") if head else None
    ho = torch.empty(B, 1, H, H, device="cuda") if head else None
    t = timeit(lambda: ops.conv3x3(x, w, out, bias=bias, act=ops.ACT_RELU, head_w=hw, head_b=0.1, head_out=ho))
    print(f"conv B={B} {H}x{H} {Cin}->{Cout} head={int(head)} store={int(store)}: {t:.1f} us  {2.0 * B * H * H * Cout * 9 * Cin / t * 1e-6:.0f} TFLOP/s", flush=True)
