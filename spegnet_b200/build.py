"""Builds libspegnet_b200_{fp16,bf16}.so in-tree with nvcc for sm_100a (``python -m spegnet_b200.build``)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")


def build(verbose: bool = False, jobs: int = 0) -> str:
    """Run the Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...); returns the .so paths.
    Cross-compiles without a GPU.  Raises on any compiler error."""
    jobs = jobs or (os.cpu_count() or 4)
    env = dict(os.environ)
    env.setdefault("PATH", "")
    if "/usr/local/cuda/bin" not in env["PATH"]:
        env["PATH"] = "/usr/local/cuda/bin:" + env["PATH"]
    proc = subprocess.run(["make", "-C", CSRC, f"-j{jobs}", "all"], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stdout.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building libspegnet_b200_*.so failed (see output above)")
    sos = [os.path.join(HERE, f"libspegnet_b200_{v}.so") for v in ("fp16", "bf16")]
    for so in sos:
        if not os.path.exists(so):
            raise RuntimeError(f"{so} was not produced")
    return sos


if __name__ == "__main__":
    print(build(verbose=True))
