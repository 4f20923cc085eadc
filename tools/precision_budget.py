"""Where the 16-bit builds lose precision, and what the library (eager PyTorch) loses on the same fixture.

    python tools/precision_budget.py [--size 512] [--out gpurun_out/precision_budget.md]

1. Per-block error table of the fp16 and the bf16 build against the fp32 oracle (stream snapshot after every trunk
   block via `SPEGNet._debug_taps` / `oracle.spegnet_forward(taps=)`, then the head tensors and the masks).
2. The oracle port itself moved to the GPU and run under `torch.autocast(bfloat16)`, `torch.autocast(float16)` and in
   fp32 with TF32 allowed: max |delta sigmoid| of the three masks against the same fp32 truth.  That is the number the
   reference would produce if its evaluator were switched to mixed precision (engine/evaluator.py:522-524 is fp32;
   engine/trainer.py:345-347 trains under fp16 autocast).
GPU box only (development / evidence tool; the oracle is used as the checker).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.init import spread_state_dict  # noqa: E402
from oracle.spegnet import spegnet_forward  # noqa: E402
from spegnet_b200 import SPEGNet  # noqa: E402

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}


def rms_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def sig_err(a, b):
    return float((a.float().cpu().sigmoid() - b.float().cpu().sigmoid()).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--out", default="gpurun_out/precision_budget.md")
    args = ap.parse_args()
    S = args.size
    sd = spread_state_dict(0)
    x = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(1))
    taps = {}
    ref = spegnet_forward(sd, x, taps=taps)  # fp32, CPU: the truth
    rows = {}
    result = {"size": S, "builds": {}, "library": {}}
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        model = SPEGNet(CFG, compute_dtype=dt, cuda_graph_max_batch=0)
        model.load_state_dict(sd)
        model = model.cuda().eval()
        model._debug_taps = {}
        out = model(x.cuda())
        torch.cuda.synchronize()
        col = {k: rms_rel(v, taps[k]) for k, v in model._debug_taps.items()}
        for k in ("fused", "context", "edge_features"):
            col["head." + k] = rms_rel(out["features"][k], ref["features"][k])
        for i in range(3):
            col[f"pred{i + 1}.logit_rms_abs"] = float((out["predictions"][i].cpu() - ref["predictions"][i]).pow(2).mean().sqrt())
            col[f"pred{i + 1}.sigmoid_max_abs"] = sig_err(out["predictions"][i], ref["predictions"][i])
        col["edge.sigmoid_max_abs"] = sig_err(out["edge"], ref["edge"])
        rows[name] = col
        result["builds"][name] = col
        del model
    # ---- the library's own reduced-precision numbers: the oracle port on the GPU under autocast
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    xg = x.cuda()
    for name, ctx in (
        ("torch_fp32_no_tf32", None),
        ("torch_fp32_tf32", "tf32"),
        ("torch_autocast_bf16", torch.bfloat16),
        ("torch_autocast_fp16", torch.float16),
    ):
        torch.backends.cuda.matmul.allow_tf32 = ctx == "tf32"
        torch.backends.cudnn.allow_tf32 = ctx == "tf32"
        if isinstance(ctx, torch.dtype):
            with torch.autocast("cuda", dtype=ctx):
                o = spegnet_forward(sd_gpu, xg)
        else:
            o = spegnet_forward(sd_gpu, xg)
        result["library"][name] = {f"pred{i + 1}.sigmoid_max_abs": sig_err(o["predictions"][i], ref["predictions"][i]) for i in range(3)}
        result["library"][name]["edge.sigmoid_max_abs"] = sig_err(o["edge"], ref["edge"])
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    keys = list(rows["fp16"].keys())
    lines = [f"Precision budget at {S}x{S}, batch 1, seed-0 spread fixture, against the fp32 CPU oracle.", "",
             "RMS relative error of the residual stream after every trunk block / of the head tensors; masks as "
             "max |delta sigmoid| (bar: 1e-2).", "", "| tensor | fp16 build | bf16 build | bf16 / fp16 |", "|---|---:|---:|---:|"]
    for k in keys:
        a, b = rows["fp16"][k], rows["bf16"][k]
        lines.append(f"| {k} | {a:.3e} | {b:.3e} | {b / max(a, 1e-30):.1f} |")
    lines += ["", "Library (oracle port on the GPU, eager PyTorch) on the same fixture, max |delta sigmoid| vs fp32 truth:", "",
              "| mode | pred1 | pred2 | pred3 | edge |", "|---|---:|---:|---:|---:|"]
    for name, c in result["library"].items():
        lines.append(f"| {name} | {c['pred1.sigmoid_max_abs']:.3e} | {c['pred2.sigmoid_max_abs']:.3e} | "
                     f"{c['pred3.sigmoid_max_abs']:.3e} | {c['edge.sigmoid_max_abs']:.3e} |")
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text + "\n")
    with open(os.path.splitext(args.out)[0] + ".json", "w") as f:
        json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
