"""Layer-by-layer comparison of the CUDA forward with the fp32 CPU oracle (development aid; GPU box).
    python tools/gpu_debug.py [S] [B] [fp16|bf16]
"""
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.init import spread_state_dict  # noqa: E402
from oracle.spegnet import spegnet_forward  # noqa: E402
from spegnet_b200 import SPEGNet, _lib  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item(), (a - b).abs().max().item(), b.abs().max().item()


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[3] if len(sys.argv) > 3 else "fp16"]
    sd = spread_state_dict(0)
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1))
    taps = {}
    t0 = time.time()
    ref = spegnet_forward(sd, x, taps=taps)
    print(f"oracle forward {time.time() - t0:.2f}s")
    model = SPEGNet({"encoder": {"config_path": "", "checkpoint_path": ""}}, compute_dtype=dt)
    model.load_state_dict(sd)
    model = model.cuda()
    model._debug_taps = {}
    out = model(x.cuda())
    torch.cuda.synchronize()
    print("launches:", _lib.launch_count())
    for k, v in model._debug_taps.items():
        r = rel(v, taps[k])
        print(f"{k:10s} rel {r[0]:.3e} abs {r[1]:.3e} (max {r[2]:.2f})")
    for k in ("fused", "context", "edge_features"):
        r = rel(out["features"][k], ref["features"][k])
        print(f"{k:14s} rel {r[0]:.3e} abs {r[1]:.3e} (max {r[2]:.2f})")
    r = rel(out["edge"], ref["edge"])
    print(f"edge logits    abs {r[1]:.3e}  sigmoid abs {(out['edge'].cpu().sigmoid() - ref['edge'].sigmoid()).abs().max():.3e}")
    for i in range(3):
        a, b = out["predictions"][i].cpu(), ref["predictions"][i]
        print(f"pred{i + 1} logits   abs {(a - b).abs().max():.3e} mean-abs {(a - b).abs().mean():.3e} std {b.std():.2f}"
              f"  sigmoid max-abs {(a.sigmoid() - b.sigmoid()).abs().max():.3e}")
    model._debug_taps = None
    for _ in range(3):
        model(x.cuda())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xg = x.cuda()
    e0.record()
    for _ in range(5):
        model(xg)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward B={B} S={S}: {e0.elapsed_time(e1) / 5:.3f} ms")


if __name__ == "__main__":
    main()
