"""Pins oracle/head.py and oracle/spegnet.py against outputs of the REFERENCE code itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the dev container)."""
import os

import numpy as np
import torch

from oracle.head import head_forward
from oracle.spegnet import spegnet_forward


def _head_inputs(seed, hw):
    g = torch.Generator().manual_seed(seed)
    f2 = torch.randn(1, 288, hw, hw, generator=g) * 2.0
    f3 = torch.randn(1, 576, hw // 2, hw // 2, generator=g) * 3.0
    f4 = torch.randn(1, 1152, hw // 4, hw // 4, generator=g) * 3.0
    return f2, f3, f4


def test_head_matches_reference_modules(golden_dir, spread_sd):
    gold = np.load(os.path.join(golden_dir, "head_small.npz"))
    f2, f3, f4 = _head_inputs(int(gold["input_seed"]), int(gold["input_hw"]))
    with torch.inference_mode():
        out = head_forward(spread_sd, [None, f2, f3, f4])
    for i, key in enumerate(("pred1", "pred2", "pred3")):
        np.testing.assert_allclose(out["predictions"][i].numpy(), gold[key], atol=2e-4, rtol=0)
    np.testing.assert_allclose(out["edge"].numpy(), gold["edge"], atol=2e-4, rtol=0)
    for key in ("fused", "context", "edge_features"):
        np.testing.assert_allclose(out["features"][key].numpy(), gold[key].astype(np.float32), atol=2e-2, rtol=2e-3)


def test_full_forward_matches_reference_class(golden_dir, spread_sd):
    gold = np.load(os.path.join(golden_dir, "full_512.npz"))
    x = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(int(gold["input_seed"])))
    out = spegnet_forward(spread_sd, x)
    assert [tuple(p.shape) for p in out["predictions"]] == [(1, 1, 128, 128), (1, 1, 256, 256), (1, 1, 512, 512)]
    assert tuple(out["edge"].shape) == (1, 1, 64, 64)
    np.testing.assert_allclose(out["predictions"][0].numpy(), gold["pred1"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(out["predictions"][1].numpy(), gold["pred2"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(out["predictions"][2].numpy(), gold["pred3"].astype(np.float32), atol=2e-2, rtol=0)
    np.testing.assert_allclose(out["edge"].numpy(), gold["edge"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(out["features"]["context"].mean(dim=(0, 2, 3)).numpy(), gold["context_mean"], atol=1e-4)
    # the fixture is not vacuous: logits are spread over several units
    assert float(out["predictions"][2].std()) > 1.5


def test_grouped_fusion_mixes_consecutive_channels(spread_sd):
    """e-ASPP's Conv2d(640,128,1,groups=128) mixes 5 CONSECUTIVE channels of the concatenation
    (SURVEY.md section 0 item 6), not one channel per branch: check with a one-hot probe."""
    import torch.nn.functional as F

    w = spread_sd["context.fusion.0.weight"]
    probe = torch.zeros(1, 640, 1, 1)
    probe[0, 130] = 1.0  # = branch 1, channel 2
    y = F.conv2d(probe, w, groups=128).flatten()
    assert int(y.nonzero().flatten()[0]) == 26 and int(y.count_nonzero()) == 1
