"""On-GPU camouflaged-object scores (csrc/metrics.cu) against the CPU oracle (oracle/sod_metrics.py, the numpy /
scipy restatement of py_sod_metrics as the reference's utils/metrics.py:142-167 calls it), through the C-ABI.

Bars: the feature transform (nearest-foreground index) is BIT-EXACT against scipy.ndimage.distance_transform_edt,
ties included; the five fp64 scores agree to 1e-9 (they differ from numpy only in the order of fp64 summation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _cases():
    rng = np.random.default_rng(7)
    out = []

    def blobs(h, w, n):
        yy, xx = np.mgrid[:h, :w]
        g = np.zeros((h, w), bool)
        for _ in range(n):
            cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
            ry, rx = rng.uniform(0.05, 0.3) * h, rng.uniform(0.05, 0.3) * w
            g |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1
        return g

    for (h, w) in [(64, 64), (96, 80), (128, 256), (100, 76), (512, 512)]:
        g = blobs(h, w, 3)
        noisy = np.clip(g * 0.7 + rng.normal(0.15, 0.2, (h, w)), 0, 1)
        out.append((f"blob_{h}x{w}", (noisy * 255).astype(np.uint8), (g * 255).astype(np.uint8)))
    h, w = 64, 48
    g = blobs(h, w, 2)
    out.append(("noise_pred", rng.integers(0, 256, (h, w)).astype(np.uint8), (g * 255).astype(np.uint8)))
    out.append(("narrow_range", rng.integers(127, 187, (h, w)).astype(np.uint8), (g * 255).astype(np.uint8)))
    out.append(("constant_pred", np.full((h, w), 93, np.uint8), (g * 255).astype(np.uint8)))
    out.append(("empty_gt", rng.integers(0, 256, (h, w)).astype(np.uint8), np.zeros((h, w), np.uint8)))
    out.append(("full_gt", rng.integers(0, 256, (h, w)).astype(np.uint8), np.full((h, w), 255, np.uint8)))
    out.append(("perfect", (g * 255).astype(np.uint8), (g * 255).astype(np.uint8)))
    sparse = (rng.random((h, w)) < 0.02)
    out.append(("sparse_gt", rng.integers(0, 256, (h, w)).astype(np.uint8), (sparse * 255).astype(np.uint8)))
    grey_gt = rng.integers(0, 256, (h, w)).astype(np.uint8)  # threshold is > 128, not >= 128
    out.append(("grey_gt", rng.integers(0, 256, (h, w)).astype(np.uint8), grey_gt))
    return out


CASES = _cases()


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from spegnet_b200 import _lib, ops as _ops

    assert _lib.load().spg_device_check() == 0
    return _ops


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_feature_transform_is_bit_exact_against_scipy(ops, name):
    from scipy.ndimage import distance_transform_edt

    _, _, gt = next(c for c in CASES if c[0] == name)
    fg = gt > 128
    nearest, stats = ops.sod_gt_prepare(torch.from_numpy(gt)[None].cuda())
    nearest = nearest[0].cpu().numpy()
    stats = stats[0].cpu().numpy()
    ys, xs = np.nonzero(fg)
    assert stats[0] == fg.sum() and stats[1] == ys.sum() and stats[2] == xs.sum()
    if not fg.any():
        assert (nearest == -1).all()
        return
    _, idx = distance_transform_edt(~fg, return_indices=True)
    want = idx[0] * gt.shape[1] + idx[1]
    assert np.array_equal(nearest, want), f"{(nearest != want).sum()} pixels pick a different nearest foreground pixel"


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_scores_match_oracle(ops, name):
    from oracle.sod_metrics import score_pair

    _, pred, gt = next(c for c in CASES if c[0] == name)
    want = score_pair(pred, gt)
    g = torch.from_numpy(gt)[None].cuda()
    nearest, stats = ops.sod_gt_prepare(g)
    got = ops.sod_scores(torch.from_numpy(pred)[None].cuda(), g, nearest, stats)[0].cpu().tolist()
    for key, value in zip(("sm", "wfm", "mae", "em", "fm"), got):
        w = want[key]
        if np.isnan(w):
            assert np.isnan(value), key
        else:
            assert abs(value - w) <= TOL, f"{name}: {key} = {value!r}, oracle {w!r}"


def test_batched_scores_equal_per_image_scores(ops):
    """Images of one batch do not interact (per-image histograms / partials), and the result is run-to-run
    deterministic (integer atomics + fixed-order fp64 sums)."""
    same = [(p, g) for n, p, g in CASES if p.shape == (64, 48)]
    pred = torch.from_numpy(np.stack([p for p, _ in same])).cuda()
    gt = torch.from_numpy(np.stack([g for _, g in same])).cuda()
    nearest, stats = ops.sod_gt_prepare(gt)
    batch = ops.sod_scores(pred, gt, nearest, stats)
    again = ops.sod_scores(pred, gt, nearest, stats)
    assert torch.equal(batch.nan_to_num(-1.0), again.nan_to_num(-1.0))
    for i in range(len(same)):
        n1, s1 = ops.sod_gt_prepare(gt[i:i + 1].contiguous())
        one = ops.sod_scores(pred[i:i + 1].contiguous(), gt[i:i + 1].contiguous(), n1, s1)
        assert torch.equal(one[0].nan_to_num(-1.0), batch[i].nan_to_num(-1.0))


def test_metrics_processor_mirrors_reference_wrapper(ops):
    """MetricsProcessor.compute_metrics(logits [B,1,H,W], [gt]) == the reference wrapper's recipe on the CPU:
    sigmoid * 255 -> truncating uint8 (utils/metrics.py:209-210), (g * 255).byte() (:220), per-sample scores, means."""
    from oracle.sod_metrics import aggregate, quantise_like_reference, score_pair
    from spegnet_b200.metrics import MetricsProcessor

    g = torch.Generator().manual_seed(11)
    B, H, W = 5, 96, 128
    logits = torch.randn(B, 1, H, W, generator=g) * 3
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    gts = [(((yy - 40 - 3 * i) ** 2 / 400.0 + (xx - 60 + 5 * i) ** 2 / 900.0) < 1).float() for i in range(B)]
    logits = logits + torch.stack(gts)[:, None] * 4 - 2
    got = MetricsProcessor().compute_metrics(logits.cuda(), [t.cuda() for t in gts])
    rows = []
    for i in range(B):
        q = quantise_like_reference(logits[i, 0].numpy())
        rows.append(score_pair(q, (gts[i] * 255).byte().numpy()))
    want = aggregate(rows)
    # the GPU sigmoid (expf) and numpy's may differ in the last bit just below an integer grey level: allow for one
    # pixel per image to truncate to the neighbouring level (score change < 1e-5), nothing more
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-5, (k, got[k], v)


def test_full_size_properties(ops):
    """BASELINE config 4 size (512x512, batch 64): size-independent properties instead of the (slow) oracle --
    a perfect prediction scores (1, 1, 0, 1), its complement scores MAE 1, and scores are invariant to the position
    of an image inside the batch."""
    B, S = 64, 512
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    gt = torch.stack([((((yy - 256 - i) / (60.0 + i)) ** 2 + ((xx - 200 - 2 * i) / 90.0) ** 2) < 1) for i in range(B)])
    gt_u8 = (gt.to(torch.uint8) * 255).cuda()
    nearest, stats = ops.sod_gt_prepare(gt_u8)
    perfect = ops.sod_scores(gt_u8, gt_u8, nearest, stats).cpu()
    assert torch.allclose(perfect[:, 0], torch.ones(B, dtype=torch.float64), atol=1e-12)   # S-alpha
    assert torch.allclose(perfect[:, 1], torch.ones(B, dtype=torch.float64), atol=1e-12)   # weighted F
    assert float(perfect[:, 2].abs().max()) == 0.0                                          # MAE
    assert torch.allclose(perfect[:, 3], torch.ones(B, dtype=torch.float64), atol=1e-5)    # E: /(N-1)
    inverse = ops.sod_scores(255 - gt_u8, gt_u8, nearest, stats).cpu()
    assert float((inverse[:, 2] - 1.0).abs().max()) == 0.0
    assert float(inverse[:, 1].abs().max()) <= 1e-12
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    pred = torch.randint(0, 256, (B, S, S), dtype=torch.uint8, device="cuda", generator=g)
    a = ops.sod_scores(pred, gt_u8, nearest, stats)
    n2, s2 = ops.sod_gt_prepare(gt_u8[perm].contiguous())
    b = ops.sod_scores(pred[perm].contiguous(), gt_u8[perm].contiguous(), n2, s2)
    assert torch.equal(a[perm], b)


def test_evaluate_dataset_matches_host_scoring_of_the_same_logits(ops, spread_sd):
    """Evaluator path end to end (engine/evaluator.py:522-560): forward -> sigmoid -> MetricsProcessor (sigmoid * 255
    -> byte again) -> five scores -> dataset means, all on the GPU, against the oracle scoring the same logits on the
    host with the reference wrapper's recipe."""
    from oracle.sod_metrics import aggregate, quantise_like_reference, score_pair
    from spegnet_b200 import SPEGNet, evaluate

    dev = torch.device("cuda", 0)
    model = SPEGNet({"encoder": {"config_path": "", "checkpoint_path": "", "variant": "large"}})
    model.load_state_dict(spread_sd)
    model = model.to(dev).eval()
    S, N = 256, 5
    fn = evaluate.synthetic_batch_fn(S, dev)
    got = evaluate.evaluate_dataset(model, N, 2, fn)  # batches of 2, 2, 1 (ragged tail)
    rows = []
    for i in range(N):
        img, gt = fn([i])
        with torch.no_grad():
            logits = model(img)["predictions"][-1][0, 0].cpu()
        q = quantise_like_reference(torch.sigmoid(logits).numpy())  # second sigmoid inside, utils/metrics.py:209
        rows.append(score_pair(q, gt[0].cpu().numpy()))
    want = aggregate(rows)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-5, (k, got[k], v)
    assert tuple(got["rows"].shape) == (N, 5)
