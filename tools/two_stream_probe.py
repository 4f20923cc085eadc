"""Does running two half-batches on two CUDA streams (HBM-bound kernels of one overlapping the tensor-bound kernels of
the other) beat one full batch?  python tools/two_stream_probe.py"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import SPEGNet  # noqa: E402

torch.manual_seed(0)
cfg = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}
m0 = SPEGNet(cfg).cuda().eval()
m1 = SPEGNet(cfg).cuda().eval()
m1.load_state_dict(m0.state_dict())
B = 64
x = torch.randn(B, 3, 512, 512, device="cuda")
xa, xb = x[: B // 2].contiguous(), x[B // 2:].contiguous()
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def one():
    with torch.no_grad():
        m0(x)


def two():
    cur = torch.cuda.current_stream()
    s0.wait_stream(cur)
    s1.wait_stream(cur)
    with torch.no_grad():
        with torch.cuda.stream(s0):
            m0(xa)
        with torch.cuda.stream(s1):
            m1(xb)
    cur.wait_stream(s0)
    cur.wait_stream(s1)


t1 = timeit(one)
t2 = timeit(two)
print(f"one stream  B=64      : {t1:.2f} ms  {B / t1 * 1e3:.0f} img/s")
print(f"two streams B=32 + 32 : {t2:.2f} ms  {B / t2 * 1e3:.0f} img/s")
