"""Stand-alone timing of the stage-1 / stage-2 GEMM shapes at batch 64 (development aid; GPU box).
    [SPG_GEMM_BRES=0] python tools/gemm_shapes_bench.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops

H = torch.float16
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

# (M, N, K, act, out fp32, residual)
shapes = [(1048576, 432, 144, 0, 0, 0), (1048576, 576, 144, 2, 0, 0), (1048576, 864, 144, 0, 0, 0), (1048576, 288, 144, 0, 1, 0),
          (1048576, 144, 168, 0, 1, 1), (1048576, 144, 144, 0, 1, 1),
          (262144, 864, 288, 0, 0, 0), (262144, 1152, 288, 2, 0, 0), (262144, 1728, 288, 0, 0, 0), (262144, 288, 288, 0, 1, 1)]
for M, N, K, act, f32, res in shapes:
    a = torch.randn(M, K, device="cuda").to(H); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(H)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else H)
    if res: out.normal_()
    t = timeit(lambda: ops.linear(a, w, out, bias=bias, act=act, residual=out if res else None))
    nbytes = M * K * 2 + N * K * 2 + M * N * (4 if f32 else 2) * (2 if res else 1)
    print(f"M={M} N={N} K={K} act={act} f32={f32} res={res}: {t:.1f} us  {2.0 * M * N * K / t * 1e-6:.0f} TFLOP/s  {nbytes / t * 1e-6:.2f} TB/s", flush=True)
