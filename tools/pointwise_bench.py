"""Times the memory-bound kernels of the head / stem alone at the batch-64 shapes, with their algorithmic bytes:
    python tools/pointwise_bench.py            (GPU box; development aid)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402

H = torch.float16
B = 64
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()  # cold L2: these kernels run once per step on data other kernels produced long ago
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def report(name, ms, nbytes):
    print(f"{name}: {ms * 1e3:.1f} us  {nbytes / ms * 1e-6:.0f} GB/s of algorithmic bytes ({nbytes / 1e6:.0f} MB)", flush=True)


# CFI combine: fp32 partial products of the three scales -> 16-bit fused map + SE row sums
Hs, C = 64, 512
g2 = torch.randn(B, Hs, Hs, C, device="cuda")
g3 = torch.randn(B, Hs // 2, Hs // 2, C, device="cuda")
g4 = torch.randn(B, Hs // 4, Hs // 4, C, device="cuda")
bias = torch.randn(C, device="cuda")
fused = torch.empty(B, Hs, Hs, C, device="cuda", dtype=H)
rs = torch.empty(B, Hs, C, device="cuda")
ms = timeit(lambda: ops.fusion_combine(g2, g3, g4, bias, fused, rs, B, Hs, C))
report("fusion_combine", ms, (g2.numel() + g3.numel() + g4.numel()) * 4 + fused.numel() * 2)

# decoder upsample + concat, stages 1 and 2
for (h, c0, he, c1, ho) in ((64, 256, 64, 64, 128), (128, 256, 64, 64, 256)):
    a = torch.randn(B, h, h, c0, device="cuda").to(H)
    e = torch.randn(B, he, he, c1, device="cuda").to(H)
    out = torch.empty(B, ho, ho, c0 + c1, device="cuda", dtype=H)
    ms = timeit(lambda: ops.upsample_concat(a, e, out))
    report(f"upsample_concat {h}->{ho}", ms, (out.numel() + a.numel() + e.numel()) * 2)
    del a, e, out

# stem im2col
x = torch.randn(B, 3, 512, 512, device="cuda")
cols = torch.empty(B * 128 * 128, 168, device="cuda", dtype=H)
ms = timeit(lambda: ops.patchify(x, cols))
report("patchify", ms, x.numel() * 4 + cols.numel() * 2)

# q-pool of the residual stream (stage 1 -> 2 geometry)
xin = torch.randn(B, 128, 128, 288, device="cuda")
y = torch.empty(B, 64, 64, 288, device="cuda")
ms = timeit(lambda: ops.maxpool2x2(xin, y, B, 128, 128, 288))
report("maxpool2x2 128->64 C=288", ms, (xin.numel() + y.numel()) * 4)

# e-ASPP branches
r128 = torch.relu(torch.randn(B, 64, 64, 128, device="cuda")).to(H)
dw = torch.randn(4, 9, 128, device="cuda") / 3
dwb = torch.randn(4, 128, device="cuda")
gvec = torch.randn(B, 128, device="cuda")
wf = torch.randn(128, 5, device="cuda")
wfb = torch.randn(128, device="cuda")
y128 = torch.empty(B, 64, 64, 128, device="cuda", dtype=H)
ms = timeit(lambda: ops.easpp_branches(r128, dw, dwb, gvec, wf, wfb, y128, B, 64, 64, (1, 6, 12, 18)))
report("easpp_branches", ms, (r128.numel() + y128.numel()) * 2)
