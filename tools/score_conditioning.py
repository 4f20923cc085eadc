"""How sensitive each of the five scores is to logit noise, on the two parity fixtures (SURVEY.md section 7 hard part 4):
the seed-0 "spread" fixture (a continuum of logits, sigma ~ 2.4) and its saturated variant (prediction heads x 8, like a
trained model's masks).  Gaussian noise of a given sigma is added to the fp32 oracle's finest logits and the scores are
recomputed through both of the reference's quantisation paths (utils/metrics.py:209-210 on logits = trainer path; on an
already sigmoided map = evaluator path).  CPU only (oracle = checker).

    python tools/score_conditioning.py [--size 256] [--batch 2] > profiles/r02_score_conditioning.md
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sod_metrics as M  # noqa: E402
from oracle.init import spread_state_dict  # noqa: E402
from oracle.spegnet import spegnet_forward  # noqa: E402


def ellipse_gt(n, size, seed=100):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    out = []
    for _ in range(n):
        m = np.zeros((size, size), bool)
        for _ in range(rng.randint(1, 4)):
            cy, cx = rng.uniform(0.25, 0.75, 2) * size
            ry, rx = rng.uniform(0.08, 0.3, 2) * size
            m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        out.append((m * 255).astype(np.uint8))
    return out


def scores(logits, gts, double_sigmoid):
    rows = []
    for lg, gt in zip(logits, gts):
        a = lg
        if double_sigmoid:
            a = 1 / (1 + np.exp(-a.astype(np.float64))).astype(np.float32)
        rows.append(M.score_pair(M.quantise_like_reference(a), gt))
    return M.aggregate(rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    sd = spread_state_dict(0)
    x = torch.randn(args.batch, 3, args.size, args.size, generator=torch.Generator().manual_seed(7 + args.size))
    gts = ellipse_gt(args.batch, args.size)
    print(f"Score conditioning: |score(logits + N(0, sigma^2)) - score(logits)|, mean over 3 noise draws; {args.batch} images, "
          f"{args.size}x{args.size}; the parity bar is 1e-3, the measured fp16 logit error is ~7e-3 rms (x8 on the saturated fixture).\n")
    for name, gain in (("spread fixture (seed 0)", 1.0), ("saturated fixture (prediction heads x 8)", 8.0)):
        sd2 = {k: v.clone() for k, v in sd.items()}
        for i in range(3):
            sd2[f"decoder.pred_heads.{i}.weight"] *= gain
            sd2[f"decoder.pred_heads.{i}.bias"] *= gain
        logits = spegnet_forward(sd2, x)["predictions"][-1][:, 0].numpy()
        print(f"### {name}: logit std {logits.std():.2f}\n")
        for ds in (False, True):
            base = scores(logits, gts, ds)
            print(f"{'evaluator path (sigmoid twice)' if ds else 'trainer path (sigmoid once)'}: "
                  + ", ".join(f"{k} {v:.4f}" for k, v in base.items()) + "\n")
            print("| logit noise sigma | s_alpha | weighted_f | mae | e_phi | mean_f |")
            print("|---:|---:|---:|---:|---:|---:|")
            for sigma in (1e-3, 3e-3, 1e-2, 3e-2, 1e-1):
                acc = {k: 0.0 for k in base}
                for draw in range(3):
                    rng = np.random.RandomState(1000 * draw + int(sigma * 1e4))
                    noisy = logits + rng.normal(0.0, sigma * gain, logits.shape).astype(np.float32)
                    s = scores(noisy, gts, ds)
                    for k in acc:
                        acc[k] += abs(s[k] - base[k]) / 3
                print(f"| {sigma:g}{' x8' if gain != 1 else ''} | " + " | ".join(f"{acc[k]:.2e}" for k in ("s_alpha", "weighted_f", "mae", "e_phi", "mean_f")) + " |")
            print()


if __name__ == "__main__":
    main()
