"""p50 batch-N latency of the drop-in forward (CUDA-graph replay for small batches): python tools/latency_b1.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import SPEGNet  # noqa: E402

torch.manual_seed(0)
model = SPEGNet({"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}).cuda().eval()
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    x = torch.randn(B, 3, 512, 512, device="cuda")
    with torch.no_grad():
        for _ in range(5):
            model(x)
        torch.cuda.synchronize()
        lat = []
        for _ in range(40):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model(x)
            b.record()
            torch.cuda.synchronize()
            lat.append(a.elapsed_time(b))
    lat.sort()
    print(f"B={B}: p50 {lat[len(lat) // 2]:.3f} ms  p90 {lat[int(len(lat) * 0.9)]:.3f} ms  ({B / lat[len(lat) // 2] * 1e3:.0f} img/s)")
