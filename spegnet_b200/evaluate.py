"""Batch-sharded dataset evaluation: the scoring loop of the reference's ``engine/evaluator.py`` with everything after
the forward kept on the GPU.

Reference flow per batch (engine/evaluator.py:522-560): ``outputs = model(images)``; per image, the finest logits are
resized to the ground truth's size, ``.sigmoid()`` is applied, and ``MetricsProcessor.compute_metrics`` (which applies
``sigmoid() * 255 -> byte`` again, utils/metrics.py:209-210) scores the pair on the host; dataset scores are plain
means over images (:447-457).  Here: forward -> ``mask_stats`` (double sigmoid + truncating uint8, one kernel) ->
``sod_scores`` (five fp64 scores per image) on the device; image ``i`` goes to rank ``i % world`` and the ranks
exchange one ``[ceil(N/world), 6]`` fp64 all-gather per dataset (``sharded.gather_rows``), after which every rank
averages in index order -- the result does not depend on the world size.

Ground truth may be a [b,S,S] tensor at the prediction's size (the synthetic COD10K-sized set of BASELINE config 4:
one batched scoring call) or a list of [h_i, w_i] masks at their original sizes, in which case every prediction is
resized to its own mask first (`spg_resize_bilinear_f32` + sigmoid), as engine/evaluator.py:539-544 does.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch

from . import metrics, ops, sharded

# (indices) -> (images fp32 [b,3,S,S] on the model's device, ground truth uint8 [b,S,S] with foreground > 128)
BatchFn = Callable[[List[int]], Tuple[torch.Tensor, torch.Tensor]]


def synthetic_pair(index: int, size: int, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Image / ground-truth pair `index` of the synthetic evaluation set: image = N(0,1)
    noise seeded 1000 + index, ground truth = union of 1-3 seeded ellipses.  Generated on the device."""
    g = torch.Generator(device=device).manual_seed(1000 + index)
    img = torch.randn(3, size, size, device=device, generator=g)
    p = torch.rand(3, 5, device=device, generator=g)
    yy = torch.arange(size, device=device, dtype=torch.float32)[:, None] / size
    xx = torch.arange(size, device=device, dtype=torch.float32)[None, :] / size
    gt = torch.zeros(size, size, dtype=torch.bool, device=device)
    for k in range(1 + index % 3):
        cy, cx = 0.2 + 0.6 * p[k, 0], 0.2 + 0.6 * p[k, 1]
        ry, rx = 0.04 + 0.2 * p[k, 2], 0.04 + 0.2 * p[k, 3]
        gt |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    return img, gt.to(torch.uint8) * 255


def synthetic_batch_fn(size: int, device: torch.device) -> BatchFn:
    def fn(indices: List[int]):
        pairs = [synthetic_pair(i, size, device) for i in indices]
        if not pairs:
            return (torch.empty(0, 3, size, size, device=device),
                    torch.empty(0, size, size, dtype=torch.uint8, device=device))
        return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])

    return fn


@torch.no_grad()
def score_batch(model, images: torch.Tensor, gt_u8: torch.Tensor) -> torch.Tensor:
    """One evaluator batch -> fp64 [b,5] rows (sm, wfm, mae, em, fm) on the device."""
    if images.shape[0] == 0:
        return torch.zeros(0, 5, dtype=torch.float64, device=images.device)
    out = model(images)
    logits = out["predictions"][-1]
    if isinstance(gt_u8, torch.Tensor) and tuple(logits.shape[-2:]) == tuple(gt_u8.shape[-2:]):
        rows, _ = metrics.per_sample(logits, metrics.PreparedGT(gt_u8), double_sigmoid=True)
        return rows
    # original-size masks: resize + sigmoid per image (engine/evaluator.py:539-544), then the wrapper's own
    # sigmoid * 255 -> byte (utils/metrics.py:209-210) and the five scores
    rows = []
    for i in range(logits.shape[0]):
        g = gt_u8[i]
        prob = ops.resize_bilinear(logits[i, 0].contiguous()[None], tuple(g.shape[-2:]), sigmoid=True)
        rows.append(metrics.per_sample(prob, metrics.PreparedGT(metrics.quantise_gt(g.reshape(1, *g.shape[-2:]))),
                                       double_sigmoid=False)[0])
    return torch.cat(rows)


@torch.no_grad()
def evaluate_dataset(model, n_items: int, batch_size: int, batch_fn: BatchFn) -> Dict[str, object]:
    """Score `n_items` images, sharded over the initialised process group (or one process).  Returns the reference's
    aggregate keys (utils/metrics.py:269-275) plus the per-image rows."""
    rows = sharded.sharded_map(n_items, batch_size, lambda idx: score_batch(model, *batch_fn(idx)))
    mean = sharded.mean_in_index_order(rows).tolist()
    out: Dict[str, object] = dict(zip(metrics.AGG_KEYS, mean))
    out["rows"] = rows
    return out
