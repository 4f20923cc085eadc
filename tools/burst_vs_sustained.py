"""Is the batch-64 step bound by the board's power limit?  Times the SAME forward (a) back to back for ~4 s and
(b) one step at a time after the GPU idled for a second (clocks at their maximum, no power history), and samples
nvidia-smi clocks / power in both phases.   python tools/burst_vs_sustained.py        (GPU box; development aid)"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spegnet_b200 import SPEGNet  # noqa: E402

CFG = {"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}}


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
            return None, None
        return statistics.median(c for c, _ in r), statistics.median(w for _, w in r)


def main():
    torch.manual_seed(0)
    model = SPEGNet(CFG, compute_dtype=torch.float16).cuda().eval()
    x = [torch.randn(64, 3, 512, 512, device="cuda") for _ in range(2)]
    smi = Smi()
    with torch.no_grad():
        for i in range(3):
            model(x[i & 1])
        torch.cuda.synchronize()
        # (a) sustained
        n = 60
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(n):
            model(x[i & 1])
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        clk, pw = smi.window(t0 + 1.0, t1)
        print(f"sustained: {e0.elapsed_time(e1) / n:.2f} ms/step over {n} steps; SM clock {clk} MHz, board power {pw} W", flush=True)
        # (b) isolated steps
        times = []
        for i in range(8):
            time.sleep(1.0)
            e0.record()
            model(x[i & 1])
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        print("isolated (1 s idle before each): " + " ".join(f"{t:.2f}" for t in times) + f" ms; median {statistics.median(times):.2f}", flush=True)
        # (c) short bursts: k steps after an idle second
        for k in (2, 4, 8, 16):
            time.sleep(1.0)
            e0.record()
            for i in range(k):
                model(x[i & 1])
            e1.record()
            torch.cuda.synchronize()
            print(f"burst of {k:2d} after 1 s idle: {e0.elapsed_time(e1) / k:.2f} ms/step", flush=True)
    smi.p.terminate()


if __name__ == "__main__":
    main()
