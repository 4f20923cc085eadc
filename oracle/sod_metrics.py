"""numpy / scipy restatement of the camouflaged-object scores used by the reference -- ORACLE only.

The reference computes S-alpha, weighted F-beta, adaptive E-phi, MAE and the mean of the F-beta curve
by calling the third-party package `py_sod_metrics` (PyPI `pysodmetrics`, version unpinned in
setup/environment.yml:26, NOT installed here and NOT in /root/reference) from
utils/metrics.py:84-86,134-167.  **Parity with that package is unpinned**: this file restates the
published algorithms (Fan et al. S-measure ICCV'17, E-measure IJCAI'18; Margolin et al. weighted
F-measure CVPR'14) as implemented by that package as far as recalled, and is pinned only by the
analytic known-answer tests in tests/test_oracle_metrics.py.

`score_pair(pred_u8, gt_u8)` mirrors utils/metrics.py:142-167 (`_process_single_sample`);
`quantise_like_reference` mirrors utils/metrics.py:205-225 (sigmoid -> *255 -> truncating uint8).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
from scipy.ndimage import convolve, distance_transform_edt

_EPS = np.spacing(1)


def prepare(pred_u8: np.ndarray, gt_u8: np.ndarray):
    """gt > 128 -> bool; pred / 255, min-max normalised when not constant."""
    gt = gt_u8 > 128
    pred = pred_u8.astype(np.float64) / 255.0
    lo, hi = pred.min(), pred.max()
    if hi != lo:
        pred = (pred - lo) / (hi - lo)
    return pred, gt


# ---------------------------------------------------------------------------------------------- MAE
def mae(pred: np.ndarray, gt: np.ndarray) -> float:
    return float(np.mean(np.abs(pred - gt)))


# ---------------------------------------------------------------------------------------- S-measure
def _s_object_part(x: np.ndarray, mask: np.ndarray) -> float:
    vals = x[mask]
    mu = vals.mean()
    sigma = vals.std(ddof=1)
    return 2.0 * mu / (mu * mu + 1.0 + sigma + _EPS)


def _s_object(pred: np.ndarray, gt: np.ndarray) -> float:
    u = gt.mean()
    fg = _s_object_part(pred * gt, gt)
    bg = _s_object_part((1.0 - pred) * (1.0 - gt), ~gt)
    return u * fg + (1.0 - u) * bg


def _centroid(gt: np.ndarray):
    h, w = gt.shape
    if np.count_nonzero(gt) == 0:
        cx, cy = np.round(w / 2), np.round(h / 2)
    else:
        cy, cx = np.argwhere(gt).mean(axis=0).round()
    return int(cx) + 1, int(cy) + 1


def _ssim(pred: np.ndarray, gt: np.ndarray) -> float:
    n = pred.size
    x, y = pred.mean(), gt.mean()
    sx = np.sum((pred - x) ** 2) / (n - 1)
    sy = np.sum((gt - y) ** 2) / (n - 1)
    sxy = np.sum((pred - x) * (gt - y)) / (n - 1)
    alpha = 4.0 * x * y * sxy
    beta = (x * x + y * y) * (sx + sy)
    if alpha != 0:
        return alpha / (beta + _EPS)
    return 1.0 if beta == 0 else 0.0


def _s_region(pred: np.ndarray, gt: np.ndarray) -> float:
    h, w = gt.shape
    x, y = _centroid(gt)
    area = h * w
    gtf = gt.astype(np.float64)
    quads = [(slice(0, y), slice(0, x)), (slice(0, y), slice(x, w)), (slice(y, h), slice(0, x)),
             (slice(y, h), slice(x, w))]
    w1 = x * y / area
    w2 = y * (w - x) / area
    w3 = (h - y) * x / area
    weights = [w1, w2, w3, 1.0 - w1 - w2 - w3]
    score = 0.0
    for wt, (sy_, sx_) in zip(weights, quads):
        p, g = pred[sy_, sx_], gtf[sy_, sx_]
        if p.size < 2:
            continue  # empty / single-pixel quadrant carries (almost) no weight
        score += wt * _ssim(p, g)
    return score


def s_measure(pred: np.ndarray, gt: np.ndarray, alpha: float = 0.5) -> float:
    y = gt.mean()
    if y == 0:
        return float(1.0 - pred.mean())
    if y == 1:
        return float(pred.mean())
    return float(max(0.0, alpha * _s_object(pred, gt) + (1.0 - alpha) * _s_region(pred, gt)))


# ------------------------------------------------------------------------------- adaptive E-measure
def adaptive_threshold(pred: np.ndarray) -> float:
    return min(2.0 * pred.mean(), 1.0)


def e_measure_adaptive(pred: np.ndarray, gt: np.ndarray, thr: float | None = None) -> float:
    """Adaptive-threshold E-measure.  `thr` overrides the adaptive threshold (tests use it to separate a real
    mask difference from the score's own discontinuity: the threshold 2*mean(pred) sits between grey levels,
    and an infinitesimal change of the mean can move one whole grey level of pixels across it)."""
    size = gt.size
    n_gt_fg = int(np.count_nonzero(gt))
    if thr is None:
        thr = adaptive_threshold(pred)
    binar = pred >= thr
    fg_fg = int(np.count_nonzero(binar & gt))
    fg_bg = int(np.count_nonzero(binar & ~gt))
    n_pred_fg = fg_fg + fg_bg
    n_pred_bg = size - n_pred_fg
    if n_gt_fg == 0:
        total = n_pred_bg
    elif n_gt_fg == size:
        total = n_pred_fg
    else:
        bg_fg = n_gt_fg - fg_fg
        bg_bg = n_pred_bg - bg_fg
        mp, mg = n_pred_fg / size, n_gt_fg / size
        combos = [(1 - mp, 1 - mg, fg_fg), (1 - mp, -mg, fg_bg), (-mp, 1 - mg, bg_fg), (-mp, -mg, bg_bg)]
        total = 0.0
        for a, b, count in combos:
            align = 2.0 * a * b / (a * a + b * b + _EPS)
            total += (align + 1.0) ** 2 / 4.0 * count
    return float(total / (size - 1 + _EPS))


# ------------------------------------------------------------------------------ weighted F-measure
def _gauss7(sigma: float = 5.0) -> np.ndarray:
    ax = np.arange(-3, 4, dtype=np.float64)
    yy, xx = np.meshgrid(ax, ax, indexing="ij")
    k = np.exp(-(xx * xx + yy * yy) / (2.0 * sigma * sigma))
    k[k < np.finfo(k.dtype).eps * k.max()] = 0
    s = k.sum()
    return k / s if s != 0 else k


def weighted_f(pred: np.ndarray, gt: np.ndarray, beta2: float = 1.0) -> float:
    if not gt.any():
        return 0.0
    dst, idx = distance_transform_edt(~gt, return_indices=True)
    err = np.abs(pred - gt.astype(np.float64))
    et = err.copy()
    bg = ~gt
    et[bg] = err[idx[0][bg], idx[1][bg]]
    ea = convolve(et, weights=_gauss7(), mode="constant", cval=0.0)
    min_e = np.where(gt & (ea < err), ea, err)
    weight = np.where(bg, 2.0 - np.exp(np.log(0.5) / 5.0 * dst), 1.0)
    ew = min_e * weight
    tpw = gt.sum() - ew[gt].sum()
    fpw = ew[bg].sum()
    recall = 1.0 - ew[gt].mean()
    precision = tpw / (tpw + fpw + _EPS)
    return float((1.0 + beta2) * recall * precision / (recall + beta2 * precision + _EPS))


# ---------------------------------------------------------------------------------- F-measure curve
def f_curve(pred: np.ndarray, gt: np.ndarray, beta2: float = 0.3) -> np.ndarray:
    q = (pred * 255).astype(np.uint8)
    bins = np.linspace(0, 256, 257)
    fg_hist, _ = np.histogram(q[gt], bins=bins)
    bg_hist, _ = np.histogram(q[~gt], bins=bins)
    tp = np.cumsum(fg_hist[::-1])
    pos = tp + np.cumsum(bg_hist[::-1])
    pos = np.where(pos == 0, 1, pos)
    total = max(int(np.count_nonzero(gt)), 1)
    precision = tp / pos
    recall = tp / total
    num = (1.0 + beta2) * precision * recall
    den = np.where(num == 0, 1.0, beta2 * precision + recall)
    return num / den


# ------------------------------------------------------------------------------------- entry points
def score_pair(pred_u8: np.ndarray, gt_u8: np.ndarray) -> Dict[str, float]:
    """One (prediction, ground truth) uint8 pair -> the five per-sample scores
    (keys as in utils/metrics.py:161-167)."""
    pred, gt = prepare(np.asarray(pred_u8), np.asarray(gt_u8))
    return {"sm": s_measure(pred, gt), "wfm": weighted_f(pred, gt), "mae": mae(pred, gt),
            "em": e_measure_adaptive(pred, gt), "fm": float(f_curve(pred, gt).mean())}


def quantise_like_reference(prob_or_logit: np.ndarray) -> np.ndarray:
    """utils/metrics.py:205-210: x.sigmoid() * 255 -> .byte() (truncation).  The evaluator hands it an
    already-sigmoided map (engine/evaluator.py:544), i.e. the sigmoid is applied twice on that path."""
    x = np.asarray(prob_or_logit, dtype=np.float32)
    return (1.0 / (1.0 + np.exp(-x)) * 255.0).astype(np.uint8)


def aggregate(rows: Sequence[Dict[str, float]]) -> Dict[str, float]:
    """utils/metrics.py:268-275 (`_aggregate_results`): plain means, reference key names."""
    n = len(rows)
    return {"s_alpha": sum(r["sm"] for r in rows) / n, "weighted_f": sum(r["wfm"] for r in rows) / n,
            "mae": sum(r["mae"] for r in rows) / n, "e_phi": sum(r["em"] for r in rows) / n,
            "mean_f": sum(r["fm"] for r in rows) / n}
