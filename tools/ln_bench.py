"""Times the residual GEMMs of the trunk with and without the producer-applied LayerNorm, and the separate LayerNorm
kernel, at the batch-64 shapes (development aid; GPU box).   python tools/ln_bench.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

H = torch.float16
shapes = [(65536, 576, 576), (65536, 576, 2304), (262144, 288, 288), (262144, 288, 1152), (1048576, 144, 144), (1048576, 144, 576)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").to(H); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(H)
    bias = torch.randn(N, device="cuda"); x = torch.randn(M, N, device="cuda"); y = torch.empty(M, N, device="cuda", dtype=H)
    g = torch.ones(N, device="cuda"); b = torch.zeros(N, device="cuda")
    t_plain = timeit(lambda: ops.linear(a, w, x, bias=bias, residual=x))
    t_ln = timeit(lambda: ops.layernorm(x, g, b, y, 1e-6))
    t_fused = timeit(lambda: ops.linear(a, w, x, bias=bias, residual=x, ln_apply=(g, b, y, 1e-6)))
    print(f"M={M} N={N} K={K}: gemm {t_plain:.1f} us + layernorm {t_ln:.1f} us = {t_plain + t_ln:.1f} | fused {t_fused:.1f} us", flush=True)
