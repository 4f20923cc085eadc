"""The reference's call sites replayed against the drop-in, step by step, from a checkpoint FILE in the trainer's
schema: `Predictor._initialize_model` + `predict_single` (engine/predictor.py:274-290, 336-368) and
`Evaluator._process_batch` (engine/evaluator.py:505-560) with `MetricsProcessor.compute_metrics`
(utils/metrics.py:169-250).  The same sequence is run on the fp32 CPU oracle and the results compared."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

MODEL_CFG = {"name": "SPEGNet",
             "encoder": {"config_path": "configs/sam2.1/sam2.1_hiera_l.yaml",
                         "checkpoint_path": "./checkpoints/sam2.1_hiera_large.pt", "variant": "large"},
             "image_processing": {"target_size": 256, "normalize_mean": [0.485, 0.456, 0.406],
                                  "normalize_std": [0.229, 0.224, 0.225]}}


@pytest.fixture(scope="module")
def checkpoint_file(tmp_path_factory, spread_sd):
    """engine/trainer.py:588-606: the dict `Trainer._save_checkpoint` writes (optimizer / scheduler / scaler states are
    opaque to the loaders; placeholders of the right kind)."""
    path = tmp_path_factory.mktemp("ckpt") / "model_best.pth"
    torch.save({"epoch": 7, "model_state_dict": spread_sd, "optimizer_state_dict": {"state": {}, "param_groups": []},
                "scheduler_state_dict": {"best": 0.5}, "scaler": {"scale": 65536.0}, "metrics": {"s_alpha": 0.5},
                "config": {"training": {"batch_size": 42}, "model": MODEL_CFG}}, path)
    return str(path)


def _load_like_the_engines(checkpoint_file, batch_size):
    from spegnet_b200 import SPEGNet

    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    device = torch.device("cuda")
    checkpoint = torch.load(checkpoint_file, map_location=device, weights_only=False)  # predictor.py:277
    model_config = checkpoint["config"]["model"]                                       # main.py:125-128
    model = SPEGNet(model_config)
    model.load_state_dict(checkpoint["model_state_dict"])                              # predictor.py:278
    model = model.to(device)
    model.eval()
    target = model_config["image_processing"]["target_size"]
    dummy = torch.randn(batch_size, 3, target, target, device=device)                  # predictor.py:283-284
    with torch.inference_mode():
        _ = model(dummy)
        torch.cuda.synchronize()                                                       # predictor.py:286-288
    return model, device, target


def test_predictor_sequence_from_checkpoint_file(checkpoint_file, spread_sd):
    from oracle.spegnet import spegnet_forward

    model, device, target = _load_like_the_engines(checkpoint_file, 1)
    x = torch.randn(1, 3, target, target, generator=torch.Generator().manual_seed(41))
    output_size = (300, 411)  # the original image size (predictor.py:350-365)
    with torch.no_grad():                                                              # predictor.py:336-338
        outputs = model(x.to(device))
        final = F.interpolate(outputs["predictions"][-1], size=output_size, mode="bilinear", align_corners=False)
        edge = F.interpolate(outputs["edge"], size=output_size, mode="bilinear", align_corners=False)
        seg_np = final.sigmoid().squeeze().cpu().numpy()                               # predictor.py:367-368
        edge_np = edge.sigmoid().squeeze().cpu().numpy()
    ref = spegnet_forward(spread_sd, x)
    want = F.interpolate(ref["predictions"][-1], size=output_size, mode="bilinear", align_corners=False).sigmoid().squeeze().numpy()
    want_e = F.interpolate(ref["edge"], size=output_size, mode="bilinear", align_corners=False).sigmoid().squeeze().numpy()
    assert seg_np.shape == output_size and seg_np.dtype == np.float32
    assert float(np.abs(seg_np - want).max()) <= 1e-2
    assert float(np.abs(edge_np - want_e).max()) <= 1e-2


def test_evaluator_batch_sequence_from_checkpoint_file(checkpoint_file, spread_sd):
    """Ragged ground-truth sizes, per-image slice -> resize -> sigmoid -> compute_metrics(seg_pred=tensor, seg_gt=[mask]),
    summed over the batch (evaluator.py:533-575); the drop-in's MetricsProcessor against the oracle's scores on the
    oracle's masks.  E-phi is held at 1e-3 only on the saturated variant (tests/test_gpu_model.py); here its bar is the
    one-grey-level bound of the spread fixture."""
    from oracle import sod_metrics as M
    from oracle.spegnet import spegnet_forward
    from spegnet_b200.metrics import MetricsProcessor

    model, device, target = _load_like_the_engines(checkpoint_file, 3)
    g = torch.Generator().manual_seed(43)
    images = torch.randn(3, 3, target, target, generator=g)
    sizes = [(256, 256), (231, 310), (400, 333)]
    rng = np.random.RandomState(5)
    masks = []
    for h, w in sizes:
        yy, xx = np.mgrid[0:h, 0:w]
        cy, cx, ry, rx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w, rng.uniform(0.15, 0.3) * h, rng.uniform(0.15, 0.3) * w
        masks.append(torch.from_numpy((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0).astype(np.float32))[None])
    proc = MetricsProcessor()
    dev_images = images.to(device)
    dev_masks = [m.to(device) for m in masks]
    torch.cuda.synchronize()                                                           # evaluator.py:507-511
    totals = {k: 0.0 for k in ("s_alpha", "weighted_f", "mae", "e_phi", "mean_f")}
    with torch.inference_mode():                                                       # evaluator.py:522-524
        outputs = model(dev_images)
        seg_preds, edge_pred = outputs["predictions"], outputs["edge"]
        for idx in range(3):
            seg_resized = F.interpolate(seg_preds[-1][idx:idx + 1], size=dev_masks[idx].shape[-2:], mode="bilinear",
                                        align_corners=False).sigmoid()                  # evaluator.py:539-544
            stage_preds = [p[idx:idx + 1].sigmoid() for p in seg_preds]                # evaluator.py:547
            assert [tuple(p.shape[-2:]) for p in stage_preds] == [(target // 4,) * 2, (target // 2,) * 2, (target,) * 2]
            F.interpolate(edge_pred[idx:idx + 1], size=dev_masks[idx].shape[-2:], mode="bilinear", align_corners=False).sigmoid()
            sample = proc.compute_metrics(seg_pred=seg_resized, seg_gt=[dev_masks[idx]])  # evaluator.py:557-560
            for k in totals:
                totals[k] += sample[k]
    ref = spegnet_forward(spread_sd, images)
    want = {k: 0.0 for k in totals}
    for idx in range(3):
        r = F.interpolate(ref["predictions"][-1][idx:idx + 1], size=masks[idx].shape[-2:], mode="bilinear",
                          align_corners=False).sigmoid()
        pred_u8 = M.quantise_like_reference(r[0, 0].numpy())  # the wrapper applies its own (second) sigmoid (metrics.py:209)
        row = M.aggregate([M.score_pair(pred_u8, (masks[idx][0].numpy() * 255).astype(np.uint8))])
        for k in want:
            want[k] += row[k]
    for k in ("s_alpha", "weighted_f", "mae", "mean_f"):
        assert abs(totals[k] - want[k]) / 3 <= 1e-3, (k, totals[k] / 3, want[k] / 3)
    assert abs(totals["e_phi"] - want["e_phi"]) / 3 <= 2.5e-2, (totals["e_phi"] / 3, want["e_phi"] / 3)
