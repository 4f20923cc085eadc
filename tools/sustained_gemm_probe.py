"""Energy view of the stage-3 GEMM shapes: each launch repeated back to back for ~2.5 s (the board's power limiter
engaged), this engine with its fused epilogue next to cuBLAS' bare matmul on the same operands, with the SM clock and
board power nvidia-smi reports in the second half of the window.  energy per launch = power x time.
    python tools/sustained_gemm_probe.py            (GPU box; development aid)"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import ops  # noqa: E402

H = torch.float16


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "50"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.p.stdout:
            try:
                c, w = (float(v) for v in line.split(","))
                self.rows.append((time.time(), c, w))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [(c, w) for (t, c, w) in self.rows if t0 <= t <= t1]
        if not r:
            return float("nan"), float("nan")
        return statistics.median(c for c, _ in r), statistics.median(w for _, w in r)


def sustained(fn, seconds=2.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # calibrate the launch count
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(20, int(seconds * 1e3 / (e0.elapsed_time(e1) / 20)))
    t0 = time.time()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    return e0.elapsed_time(e1) / n * 1e3, t0, t1  # us per launch


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""   # optional filter on the launch name, e.g. "fc2"
    smi = Smi()
    M = 65536
    print("| launch | us (sustained) | TFLOP/s | SM MHz | board W | mJ per launch | pJ per FLOP |\n|---|---:|---:|---:|---:|---:|---:|")
    for name, N, K, kind in (("qkv", 1728, 576, "plain"), ("fc1+GELU", 2304, 576, "gelu"), ("proj+res", 576, 576, "res"),
                             ("fc2+res", 576, 2304, "res"), ("fc2+res+LN", 576, 2304, "ln"), ("square 2304", 2304, 2304, "plain")):
        if only and only not in name:
            continue
        a = torch.randn(M, K, device="cuda").to(H)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(H)
        bias = torch.randn(N, device="cuda")
        fl = 2.0 * M * N * K
        rows = []
        if kind in ("plain", "gelu"):
            out = torch.empty(M, N, device="cuda", dtype=H)
            act = ops.ACT_GELU if kind == "gelu" else ops.ACT_NONE
            rows.append(("engine " + name, lambda: ops.linear(a, w, out, bias=bias, act=act)))
            if kind == "gelu":  # the same launch without the activation: what the GELU itself costs
                rows.append(("engine " + name + " minus GELU", lambda: ops.linear(a, w, out, bias=bias)))
        else:
            out16 = torch.empty(M, N, device="cuda", dtype=H)
            rows.append(("engine " + name + " minus residual (16-bit store)", lambda: ops.linear(a, w, out16, bias=bias)))
            x = torch.randn(M, N, device="cuda")
            if kind == "ln":
                g, b = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
                y = torch.empty(M, N, device="cuda", dtype=H)
                rows.append(("engine " + name, lambda: ops.linear(a, w, x, bias=bias, residual=x, ln_apply=(g, b, y, 1e-6))))
            else:
                rows.append(("engine " + name, lambda: ops.linear(a, w, x, bias=bias, residual=x)))
        if kind != "ln":
            wt = w.t()
            rows.append((f"cuBLAS matmul {N}x{K} (no epilogue)", lambda: torch.matmul(a, wt)))
        for label, fn in rows:
            us, t0, t1 = sustained(fn)
            clk, pw = smi.window(t0 + (t1 - t0) / 2, t1)
            print(f"| {label} | {us:.1f} | {fl / us * 1e-6:.0f} | {clk:.0f} | {pw:.0f} | {pw * us * 1e-3:.1f} | {pw * us * 1e-6 / fl * 1e12:.2f} |", flush=True)
            time.sleep(1.0)
    smi.p.terminate()


if __name__ == "__main__":
    main()
