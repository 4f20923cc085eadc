"""Input / output side on the GPU (csrc/imageio.cu) through the C-ABI: preprocessing against the reference
CODImageProcessor's own outputs (tests/golden/preprocess.npz) and against the oracle on fresh sizes; prediction resize
against F.interpolate; the evaluator's original-size scoring path against the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")
# fp32 in both implementations; they differ in summation order of the <= 2*scale+1 filter taps and in the fused
# (x - mean) / std rounding: a few ulp of an O(1) value
TOL = 2e-6


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spegnet_b200 import _lib, ops as _ops

    assert _lib.load().spg_device_check() == 0
    return _ops


@pytest.mark.parametrize("name", ["down", "down_big", "up", "mixed", "same_w"])
def test_preprocess_matches_reference_golden(ops, name):
    g = np.load(GOLD)
    got = ops.preprocess_rgb(torch.from_numpy(g[name + "_rgb"]).cuda(), int(g[name + "_target"])).cpu().numpy()
    want = g[name + "_out"]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= TOL * 4.5, np.abs(got - want).max()  # 1 / std amplifies by up to 4.46


@pytest.mark.parametrize("h,w,target", [(683, 1024, 512), (512, 512, 512), (1500, 2000, 512), (300, 400, 512),
                                        (768, 1024, 1024)])
def test_preprocess_matches_oracle_at_real_sizes(ops, h, w, target):
    from oracle.preprocess import process_image_array

    rng = np.random.default_rng(h * 7 + w)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = ops.preprocess_rgb(torch.from_numpy(rgb).cuda(), target).cpu()
    want = process_image_array(rgb, target)
    assert float((got - want).abs().max()) <= TOL * 4.5


@pytest.mark.parametrize("hi,ho,wo,sig", [(512, 683, 1024, True), (512, 512, 512, False), (64, 301, 203, True),
                                          (512, 300, 400, True), (128, 97, 131, False)])
def test_resize_bilinear_matches_interpolate(ops, hi, ho, wo, sig):
    g = torch.Generator(device="cuda").manual_seed(hi + ho)
    x = torch.randn(3, 1, hi, hi, device="cuda", generator=g) * 4
    got = ops.resize_bilinear(x, (ho, wo), sigmoid=sig)
    want = F.interpolate(x, size=(ho, wo), mode="bilinear", align_corners=False)
    want = want.sigmoid() if sig else want
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 5e-6


def test_original_size_scoring_matches_oracle(ops, spread_sd):
    """engine/evaluator.py:539-560 with masks at their original sizes: resize -> sigmoid -> MetricsProcessor."""
    from oracle.preprocess import resize_logits
    from oracle.sod_metrics import quantise_like_reference, score_pair
    from spegnet_b200 import SPEGNet, evaluate

    dev = torch.device("cuda", 0)
    model = SPEGNet({"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}})
    model.load_state_dict(spread_sd)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(4)
    images = torch.randn(2, 3, 256, 256, generator=g).to(dev)
    sizes = [(301, 203), (180, 333)]
    gts = []
    for h, w in sizes:
        yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        gts.append(((((yy - h / 2) / (h / 4)) ** 2 + ((xx - w / 2.5) / (w / 5)) ** 2) < 1).to(torch.uint8).mul(255).to(dev))
    rows = evaluate.score_batch(model, images, gts).cpu()
    with torch.no_grad():
        logits = model(images)["predictions"][-1].cpu()
    for i, (h, w) in enumerate(sizes):
        prob = resize_logits(logits[i:i + 1], (h, w), sigmoid=True)[0, 0].numpy()
        want = score_pair(quantise_like_reference(prob), gts[i].cpu().numpy())
        for k, key in enumerate(("sm", "wfm", "mae", "em", "fm")):
            assert abs(float(rows[i, k]) - want[key]) <= 2e-4, (i, key, float(rows[i, k]), want[key])


def test_predict_single_matches_oracle_pipeline(ops, spread_sd):
    """Predictor.predict_single end to end (engine/predictor.py:311-368): preprocess -> forward -> resize -> sigmoid on
    the GPU against the fp32 CPU oracle of every stage; masks within the north star's 1e-2."""
    from oracle.preprocess import process_image_array, resize_logits
    from oracle.spegnet import spegnet_forward
    from spegnet_b200 import SPEGNet, predict

    dev = torch.device("cuda", 0)
    model = SPEGNet({"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}})
    model.load_state_dict(spread_sd)
    model = model.to(dev).eval()
    rng = np.random.default_rng(3)
    h, w = 300, 420
    yy, xx = np.mgrid[:h, :w]
    rgb = np.clip(127 + 90 * np.sin(xx / 17.0)[..., None] * np.array([1, .5, -1]) + 60 * np.cos(yy / 23.0)[..., None]
                  + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)
    seg, edge = predict.predict_single(model, torch.from_numpy(rgb).to(dev), target_size=256, output_size=(h, w))
    x = process_image_array(rgb, 256)[None]
    ref = spegnet_forward(spread_sd, x)
    want_seg = resize_logits(ref["predictions"][-1], (h, w))[0, 0]
    want_edge = resize_logits(ref["edge"], (h, w))[0, 0]
    assert tuple(seg.shape) == (h, w) and tuple(edge.shape) == (h, w)
    assert float((seg.cpu() - want_seg).abs().max()) <= 1e-2
    assert float((edge.cpu() - want_edge).abs().max()) <= 1e-2
    mask = predict.binary_mask_u8(seg)
    assert mask.dtype == torch.uint8 and int((mask.cpu().int() - (want_seg * 255).to(torch.uint8).int()).abs().max()) <= 3


def test_collate_decoded_ragged_batch_matches_the_reference_path():
    """Ragged decoded images + ragged masks -> one device batch: every slot equals the oracle's process_image of that
    image, masks are the reference's (mask > 127.5).float() at their own sizes, and the evaluator scoring loop accepts
    the result as is."""
    import numpy as np

    from oracle.preprocess import process_image_array
    from spegnet_b200.batching import collate_decoded

    rng = np.random.RandomState(3)
    sizes = [(300, 400), (512, 512), (641, 333)]
    imgs = [rng.randint(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    masks = [rng.randint(0, 256, (h, w), dtype=np.uint8) for h, w in sizes]
    out = collate_decoded([torch.from_numpy(a).cuda() for a in imgs], [torch.from_numpy(m).cuda() for m in masks],
                          names=["a", "b", "c"], target_size=256)
    assert tuple(out["images"].shape) == (3, 3, 256, 256) and out["names"] == ["a", "b", "c"]
    for i, a in enumerate(imgs):
        want = process_image_array(a, 256)
        assert float((out["images"][i].cpu() - want).abs().max()) <= 2e-5
        ref_mask = (torch.from_numpy(masks[i]).float() > 127.5).float()[None]
        assert torch.equal(out["masks"][i].cpu(), ref_mask)
    with pytest.raises(ValueError):
        collate_decoded([])
