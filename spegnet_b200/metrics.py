"""GPU mirror of the reference's ``utils/metrics.py::MetricsProcessor``.

Same public surface -- ``MetricsProcessor(num_processes=None).compute_metrics(seg_pred, seg_gt, edge_pred=None,
edge_gt=None) -> {'s_alpha', 'weighted_f', 'mae', 'e_phi', 'mean_f'[, 'edge_mae', 'edge_f']}`` (utils/metrics.py:
169-250, 252-285) -- but the per-sample scores (utils/metrics.py:142-167, py_sod_metrics) are computed by the
sm_100a kernels of ``csrc/metrics.cu`` on the device the predictions live on: no device->host copy of the masks,
no process pool.  ``per_sample`` exposes the [B,5] fp64 rows that ``spegnet_b200.sharded`` gathers across ranks.
There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import ops

SCORE_KEYS = ("sm", "wfm", "mae", "em", "fm")                               # utils/metrics.py:161-167
AGG_KEYS = ("s_alpha", "weighted_f", "mae", "e_phi", "mean_f")              # utils/metrics.py:269-275


def quantise_gt(gt: torch.Tensor) -> torch.Tensor:
    """``(g * 255).byte()`` of utils/metrics.py:220.  The reference applies it to every mask it is given ({0,1} floats
    from CODImageProcessor.process_mask); a uint8 mask is taken as already quantised ({0,255}) unless its maximum is
    <= 1, in which case it is the same {0,1} mask in integer form and is scaled like the reference would (no host
    synchronisation: the scale factor is computed on the tensor's device)."""
    if gt.dtype == torch.uint8:
        if gt.numel() == 0:
            return gt.contiguous()
        scale = (gt.max() <= 1).to(torch.uint8) * 254 + 1
        return (gt * scale).contiguous()
    return (gt * 255).to(torch.uint8).contiguous()


class PreparedGT:
    """Ground truth uploaded and analysed once (feature transform + centroid sums): reusable across evaluations."""

    def __init__(self, gt_u8: torch.Tensor):
        if gt_u8.dim() != 3 or gt_u8.dtype != torch.uint8:
            raise ValueError("PreparedGT needs a uint8 [B,H,W] tensor")
        if not gt_u8.is_cuda:
            raise RuntimeError("spegnet_b200.metrics runs on a CUDA (B200) device only; there is no CPU fallback")
        self.gt = gt_u8.contiguous()
        self.nearest, self.stats = ops.sod_gt_prepare(self.gt)


def per_sample(logits_or_prob: torch.Tensor, gt: Union[torch.Tensor, PreparedGT], double_sigmoid: bool = False
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B,1,H,W] / [B,H,W] fp32 maps + [B,H,W] ground truth -> (scores fp64 [B,5] in SCORE_KEYS order, uint8 masks).
    The map goes through ``sigmoid() * 255 -> byte`` exactly as utils/metrics.py:209-210 does (twice with
    `double_sigmoid`, the evaluator path engine/evaluator.py:544 + utils/metrics.py:209)."""
    x = logits_or_prob
    if x.dim() == 4:
        if x.shape[1] != 1:
            raise ValueError(f"expected [B,1,H,W], got {tuple(x.shape)}")
        x = x[:, 0]
    if x.dim() != 3:
        raise ValueError(f"expected [B,1,H,W] or [B,H,W], got {tuple(logits_or_prob.shape)}")
    if not x.is_cuda:
        raise RuntimeError("spegnet_b200.metrics runs on a CUDA (B200) device only; there is no CPU fallback")
    prep = gt if isinstance(gt, PreparedGT) else PreparedGT(quantise_gt(gt))
    if tuple(prep.gt.shape) != tuple(x.shape):
        raise ValueError(f"prediction {tuple(x.shape)} and ground truth {tuple(prep.gt.shape)} differ in shape")
    mask, _ = ops.mask_stats(x.contiguous().float(), prep.gt, double_sigmoid)
    return ops.sod_scores(mask, prep.gt, prep.nearest, prep.stats), mask


def aggregate(rows: torch.Tensor) -> Dict[str, float]:
    """utils/metrics.py:252-275: plain means over samples, reference key names."""
    mean = rows.to(torch.float64).mean(dim=0).tolist()
    return dict(zip(AGG_KEYS, mean))


class MetricsProcessor:
    """Drop-in for utils/metrics.py::MetricsProcessor (constructor argument kept, unused: nothing runs on the host)."""

    def __init__(self, num_processes: Optional[int] = None):
        self.num_processes = num_processes

    @staticmethod
    def _rows(pred: Union[Sequence[torch.Tensor], torch.Tensor], gt: Sequence[torch.Tensor]) -> torch.Tensor:
        if isinstance(pred, torch.Tensor):
            same = all(tuple(g.shape[-2:]) == tuple(pred.shape[-2:]) for g in gt)
            if same:
                g = torch.stack([quantise_gt(t.reshape(t.shape[-2:])) for t in gt]).to(pred.device)
                return per_sample(pred, g)[0]
            pred = list(pred.split(1, dim=0))
        rows: List[torch.Tensor] = []
        for p, g in zip(pred, gt):  # ragged sizes (the evaluator resizes every map to its own ground truth)
            p3 = p.reshape(1, *p.shape[-2:])
            g3 = quantise_gt(g.reshape(1, *g.shape[-2:])).to(p.device)
            rows.append(per_sample(p3, g3)[0])
        return torch.cat(rows)

    @torch.no_grad()
    def compute_metrics(self, seg_pred, seg_gt, edge_pred=None, edge_gt=None) -> Dict[str, float]:
        out = aggregate(self._rows(seg_pred, seg_gt))
        if edge_pred is not None and edge_gt is not None:
            e = self._rows(edge_pred, edge_gt).mean(dim=0).tolist()
            out.update({"edge_mae": e[2], "edge_f": e[4]})
        return out
