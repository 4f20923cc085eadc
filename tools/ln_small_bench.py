"""Latency-regime costs of the three ways to get y = LayerNorm(fc2 output): GEMM + layernorm, GEMM + layernorm_matched,
fused ln_apply -- under CUDA-graph replay of 16 back-to-back repetitions (development aid; GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops

H = torch.float16
def graph_time(fn, reps=16, iters=20):
    ops.set_pdl(True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / reps * 1e3

for M, N, K in [(1024, 576, 2304), (4096, 576, 2304), (8192, 576, 2304), (16384, 576, 2304), (4096, 288, 1152), (16384, 144, 576)]:
    a = torch.randn(M, K, device="cuda").to(H); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(H)
    bias = torch.randn(N, device="cuda"); x = torch.randn(M, N, device="cuda"); y = torch.empty(M, N, device="cuda", dtype=H)
    g = torch.ones(N, device="cuda"); b = torch.zeros(N, device="cuda")
    t_g = graph_time(lambda: ops.linear(a, w, x, bias=bias, residual=x))
    t_l = graph_time(lambda: ops.layernorm(x, g, b, y, 1e-6))
    t_m = graph_time(lambda: ops.layernorm_matched(x, g, b, y, 1e-6))
    t_f = graph_time(lambda: ops.linear(a, w, x, bias=bias, residual=x, ln_apply=(g, b, y, 1e-6)))
    print(f"M={M} N={N} K={K}: gemm {t_g:.1f} | layernorm {t_l:.1f} | matched {t_m:.1f} | fused gemm+ln {t_f:.1f} us", flush=True)
