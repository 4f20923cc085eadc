"""In-step per-op timing of the B=64 forward with CUDA events around every C-ABI call (real clocks, warm L2, the
launch order of the model) -- complements the ncu launch list, whose per-launch times are cold-cache and serialised.

    python tools/op_profile.py [--batch 64] [--size 512] [--steps 5] [--out gpurun_out/op_profile.md]
"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from spegnet_b200 import SPEGNet, ops  # noqa: E402

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}
NAMES = ["patchify", "linear", "layernorm", "maxpool2x2", "window_attention", "cast_h16", "fusion_combine", "pooled_mlp",
         "scale_channels", "row_sums", "easpp_branches", "conv3x3", "upsample_concat", "mask_stats", "conv3x3_up2", "up2_border_gather"]


def sig(name, a, kw):
    if name == "linear":
        x, w, out = a[0], a[1], a[2]
        tag = ("gelu" if kw.get("act") == ops.ACT_GELU else "relu" if kw.get("act") == ops.ACT_RELU else "none")
        return f"linear M={x.shape[0]} N={w.shape[0]} K={w.shape[1]} act={tag} res={int(kw.get('residual') is not None)} " \
               f"out={'f32' if out.dtype == torch.float32 else 'h16'}", 2.0 * x.shape[0] * w.shape[0] * w.shape[1]
    if name == "conv3x3":
        x, w = a[0], a[1]
        return f"conv3x3 {tuple(x.shape)} -> {w.shape[0]} head={int(kw.get('head_w') is not None)}", \
            2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * w.shape[1]
    if name == "conv3x3_up2":
        x, w = a[0], a[1]
        return f"conv3x3_up2 {tuple(x.shape)} -> 4x{w.shape[0] // 12}", 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * (w.shape[0] // 3) * w.shape[1]
    if name == "window_attention":
        return f"attention B,H,W,D,heads,win,pool={a[2:]}", 0.0
    if name == "layernorm":
        return f"layernorm {tuple(a[0].shape)}", 0.0
    if name == "upsample_concat":
        return f"upsample_concat -> {tuple(a[2].shape)}", 0.0
    return name, 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_profile.md"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = SPEGNet(CFG).to(dev).eval()
    x = [torch.randn(args.batch, 3, args.size, args.size, device=dev) for _ in range(2)]
    with torch.no_grad():
        for i in range(3):
            model(x[i % 2])
        torch.cuda.synchronize()
        events = []
        for name in NAMES:
            real = getattr(ops, name)

            def wrapped(*a, _real=real, _name=name, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = _real(*a, **kw)
                e1.record()
                events.append((_name, a, kw, e0, e1))
                return r

            setattr(ops, name, wrapped)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            model(x[i % 2])
        t1.record()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for name, a, kw, e0, e1 in events:
        s, fl = sig(name, a, kw)
        r = agg[s]
        r[0] += 1
        r[1] += e0.elapsed_time(e1)
        r[2] += fl
    total = t0.elapsed_time(t1) / args.steps
    lines = [f"step {total:.3f} ms (B={args.batch}, S={args.size}, {args.steps} steps, events around every op)", "",
             "| op | launches/step | ms/step | share | us/launch | TFLOP/s |", "|---|---:|---:|---:|---:|---:|"]
    for s, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tf = f"{fl / (ms * 1e-3) * 1e-12:.0f}" if fl else ""
        lines.append(f"| {s} | {n // args.steps} | {ms / args.steps:.3f} | {100 * ms / args.steps / total:.1f} % | "
                     f"{1e3 * ms / n:.1f} | {tf} |")
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
