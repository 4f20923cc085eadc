"""Golden vectors for the input side: the REFERENCE CODImageProcessor (utils/image_processor.py), imported verbatim
from /root/reference, run on small synthetic PNG files.  Dev container only:

    python tests/golden/make_golden_preprocess.py

Writes tests/golden/preprocess.npz: for each case the decoded RGB array (uint8, what PIL hands the reference) and the
reference's processed tensor.  Cases cover down-scaling by non-integer factors in both axes, up-scaling, a mixed
case (one axis up, one down) and an axis whose size already equals the target."""
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from utils.image_processor import CODImageProcessor  # noqa: E402

CASES = [("down", 97, 131, 64), ("down_big", 301, 203, 64), ("up", 40, 50, 64), ("mixed", 48, 150, 64), ("same_w", 100, 64, 64)]


def main():
    rng = np.random.default_rng(2026)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, h, w, target in CASES:
            yy, xx = np.mgrid[:h, :w]
            base = 127 + 100 * np.sin(xx / 7.0 + yy / 11.0)[..., None] * np.array([1.0, 0.6, -0.8])
            rgb = np.clip(base + rng.normal(0, 25, (h, w, 3)), 0, 255).astype(np.uint8)
            path = os.path.join(tmp, name + ".png")
            Image.fromarray(rgb).save(path)
            proc = CODImageProcessor(target_size=target)
            ref = proc.process_image(path).numpy()
            out[name + "_rgb"] = np.array(Image.open(path).convert("RGB"))
            out[name + "_out"] = ref.astype(np.float32)
            out[name + "_target"] = np.int64(target)
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("wrote preprocess.npz", {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
