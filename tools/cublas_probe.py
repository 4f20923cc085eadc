"""Library baseline for the GEMM engine: cuBLAS (torch.matmul / F.linear, fp16, fp32 accumulate) on the exact stage-3
shapes of the batch-64 step, without and with the unfused epilogue the reference would run (bias + GELU / residual add
as separate ATen kernels).  python tools/cublas_probe.py"""
import torch
import torch.nn.functional as F


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us


M = 65536
for name, N, K, epi in (("qkv", 1728, 576, "bias"), ("proj", 576, 576, "res"), ("fc1", 2304, 576, "gelu"), ("fc2", 576, 2304, "res")):
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) / K ** 0.5
    b = torch.randn(N, device="cuda", dtype=torch.float16)
    res = torch.randn(M, N, device="cuda", dtype=torch.float32)
    t_mm = timeit(lambda: torch.matmul(x, w.t()))
    if epi == "bias":
        t_full = timeit(lambda: F.linear(x, w, b))
    elif epi == "gelu":
        t_full = timeit(lambda: F.gelu(F.linear(x, w, b)))
    else:
        t_full = timeit(lambda: res.add_(F.linear(x, w, b)))
    fl = 2.0 * M * N * K
    print(f"{name:5s} M={M} N={N} K={K}: matmul {t_mm:7.1f} us {fl / t_mm * 1e-6:7.1f} TFLOP/s | with unfused {epi:4s} epilogue {t_full:7.1f} us "
          f"{fl / t_full * 1e-6:7.1f} TFLOP/s")
