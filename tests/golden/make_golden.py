"""Generate the golden vectors under tests/golden/ by running the REFERENCE code itself.

Run in the dev container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

* head_small.npz  -- the reference's own AdaptiveAttentionFusion / EfficientASPP / EdgeDetectionModule /
                     BoundaryAwareDecoder (models/feature_integration.py, models/object_detection.py),
                     imported verbatim, on seeded multi-scale features at a reduced spatial size.
* full_512.npz    -- the reference `SPEGNet` class (models/spegnet.py) constructed verbatim around the
                     sam2 shim (oracle/sam2_shim.py; the trunk itself is third-party and absent, so that
                     part is the restated trunk), spread init seed 0, image = randn seed 1, S=512, B=1.

Weights are never stored: they are regenerated from oracle.init.spread_state_dict(seed) (deterministic
per-key CPU generators), only inputs' seeds and the reference outputs are committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import sam2_shim  # noqa: E402
from oracle.init import spread_state_dict  # noqa: E402


def head_inputs(seed: int, hw: int):
    g = torch.Generator().manual_seed(seed)
    f2 = torch.randn(1, 288, hw, hw, generator=g) * 2.0
    f3 = torch.randn(1, 576, hw // 2, hw // 2, generator=g) * 3.0
    f4 = torch.randn(1, 1152, hw // 4, hw // 4, generator=g) * 3.0
    return f2, f3, f4


def make_head_small():
    from models.feature_integration import AdaptiveAttentionFusion, EfficientASPP
    from models.object_detection import BoundaryAwareDecoder, EdgeDetectionModule

    sd = spread_state_dict(0)
    fusion = AdaptiveAttentionFusion([288, 576, 1152], 512).eval()
    context = EfficientASPP(512, 256, 4, [1, 6, 12, 18]).eval()
    edge = EdgeDetectionModule(256, 64).eval()
    dec = BoundaryAwareDecoder(256, [256, 128, 64], 1, [64, 64, None]).eval()
    for name, mod in (("fusion.", fusion), ("context.", context), ("edge_detector.", edge), ("decoder.", dec)):
        mod.load_state_dict({k[len(name):]: v for k, v in sd.items() if k.startswith(name)})
    f2, f3, f4 = head_inputs(7, 16)
    with torch.inference_mode():
        fused = fusion([f2, f3, f4])
        ctx = context(fused)
        edge_map, edge_feat = edge(ctx)
        preds = dec(ctx, edge_features_list=[edge_feat, edge_feat, None])
    np.savez_compressed(
        os.path.join(HERE, "head_small.npz"),
        input_seed=7, input_hw=16, weight_seed=0,
        pred1=preds[0].numpy(), pred2=preds[1].numpy(), pred3=preds[2].numpy(), edge=edge_map.numpy(),
        fused=fused.numpy().astype(np.float16), context=ctx.numpy().astype(np.float16),
        edge_features=edge_feat.numpy().astype(np.float16),
    )
    print("head_small:", [tuple(p.shape) for p in preds], float(preds[2].std()))


def make_full_512():
    sam2_shim.install()
    from models.spegnet import SPEGNet

    model = SPEGNet({"encoder": {"config_path": "configs/sam2.1/sam2.1_hiera_l.yaml",
                                 "checkpoint_path": "./checkpoints/sam2.1_hiera_large.pt", "variant": "large"}}).eval()
    model.load_state_dict(spread_state_dict(0))
    x = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(1))
    with torch.inference_mode():
        out = model(x)
    p = out["predictions"]
    np.savez_compressed(
        os.path.join(HERE, "full_512.npz"),
        input_seed=1, weight_seed=0,
        pred1=p[0].numpy(), pred2=p[1].numpy(), pred3=p[2].numpy().astype(np.float16), edge=out["edge"].numpy(),
        context_mean=out["features"]["context"].mean(dim=(0, 2, 3)).numpy(),
        fused_mean=out["features"]["fused"].mean(dim=(0, 2, 3)).numpy(),
    )
    print("full_512:", [tuple(t.shape) for t in p], float(p[2].std()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    make_head_small()
    make_full_512()
