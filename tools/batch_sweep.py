"""Images/s of the eager forward as a function of the batch size (device-timed, CUDA events): does a smaller batch --
a working set that stays in the 126 MB L2 between producer and consumer kernels -- buy more than its tail effects cost?

    python tools/batch_sweep.py [--size 512] [--batches 4,8,16,32,64] [--dtype fp16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import SPEGNet  # noqa: E402

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batches", default="4,8,16,32,64")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--images", type=int, default=640, help="images per timed measurement")
    args = ap.parse_args()
    torch.manual_seed(0)
    model = SPEGNet(CFG, compute_dtype=torch.float16 if args.dtype == "fp16" else torch.bfloat16, cuda_graph_max_batch=0)
    model = model.cuda().eval()
    out = {}
    for B in [int(b) for b in args.batches.split(",")]:
        xs = [torch.randn(B, 3, args.size, args.size, device="cuda") for _ in range(3)]
        with torch.no_grad():
            for i in range(3):
                model(xs[i % 3])
            torch.cuda.synchronize()
            steps = max(3, args.images // B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                model(xs[i % 3])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[B] = {"ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1)}
        print(B, out[B], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
