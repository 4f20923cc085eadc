"""One batch-64 forward inside a cudaProfilerStart/Stop range, after warm-up (for `ncu --profile-from-start off`).
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python tools/one_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import SPEGNet  # noqa: E402

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dtype", default="fp16")
    args = ap.parse_args()
    torch.manual_seed(0)
    model = SPEGNet(CFG, compute_dtype=torch.float16 if args.dtype == "fp16" else torch.bfloat16).cuda().eval()
    x = [torch.randn(args.batch, 3, args.size, args.size, device="cuda") for _ in range(2)]
    with torch.no_grad():
        model(x[0])
        model(x[1])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model(x[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
