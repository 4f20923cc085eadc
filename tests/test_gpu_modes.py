"""Non-default kernel modes that are selected by environment variables read when the library is loaded: each runs a
slice of the GPU suite in a fresh interpreter with the variable set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODES = [
    # A chunks TMA-multicast across the n-tile cluster of the LayerNorm-producing GEMM (gemm_tcgen05.cu: a_mcast)
    ({"SPG_GEMM_AMCAST": "1"}, "producer or ln_apply or layernorm"),
    # no resident weight tiles, no producer LayerNorm: the plain ring for every launch
    ({"SPG_GEMM_BRES": "0", "SPG_LN_APPLY": "0"}, "linear or patch_embed"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("env,expr", MODES, ids=lambda v: "-".join(f"{k}={x}" for k, x in v.items()) if isinstance(v, dict) else None)
def test_mode_subset(env, expr):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_ops.py"), "-x", "-q", "-m", "gpu",
                        "-k", expr, "-p", "no:cacheprovider"], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout, tail
