"""Analytic known-answer tests for the restated py_sod_metrics scores (oracle/sod_metrics.py).
The upstream package is absent and unpinned, so these closed-form cases are the only pin."""
import numpy as np
import pytest

from oracle import sod_metrics as M


def _square(n=64, lo=16, hi=48):
    gt = np.zeros((n, n), np.uint8)
    gt[lo:hi, lo:hi] = 255
    return gt


def test_perfect_prediction():
    gt = _square()
    s = M.score_pair(gt.copy(), gt)
    assert s["mae"] == 0.0
    assert s["sm"] == pytest.approx(1.0, abs=1e-9)
    assert s["wfm"] == pytest.approx(1.0, abs=1e-9)
    assert s["em"] == pytest.approx(gt.size / (gt.size - 1), rel=1e-9)  # the package's /(N-1) normalisation
    # F-curve: thresholds 1..255 give P=R=1, threshold 0 gives P=|fg|/N -> mean slightly below 1
    p0 = 0.25
    f0 = 1.3 * p0 * 1.0 / (0.3 * p0 + 1.0)
    assert s["fm"] == pytest.approx((255 + f0) / 256, rel=1e-9)


def test_inverted_prediction():
    gt = _square()
    s = M.score_pair(255 - gt, gt)
    assert s["mae"] == 1.0
    assert s["sm"] == pytest.approx(0.0, abs=1e-12)  # clamped at zero
    assert s["wfm"] == pytest.approx(0.0, abs=1e-9)
    assert s["em"] == pytest.approx(0.0, abs=1e-9)


def test_empty_ground_truth():
    gt = np.zeros((32, 32), np.uint8)
    pred = np.full((32, 32), 51, np.uint8)  # constant 0.2, not min-max normalised
    s = M.score_pair(pred, gt)
    assert s["mae"] == pytest.approx(0.2)
    assert s["sm"] == pytest.approx(0.8)
    assert s["wfm"] == 0.0
    # adaptive threshold 0.4 > 0.2 -> nothing predicted -> all pixels aligned
    assert s["em"] == pytest.approx(1024 / 1023)


def test_full_ground_truth():
    gt = np.full((32, 32), 255, np.uint8)
    pred = np.full((32, 32), 204, np.uint8)  # constant 0.8
    s = M.score_pair(pred, gt)
    assert s["sm"] == pytest.approx(0.8)
    assert s["mae"] == pytest.approx(0.2)


def test_min_max_normalisation_and_threshold():
    gt = _square()
    pred = np.where(gt > 0, 140, 120).astype(np.uint8)  # two grey levels -> normalised to {0,1}
    s = M.score_pair(pred, gt)
    assert s["mae"] == 0.0 and s["sm"] == pytest.approx(1.0)
    assert (M.prepare(np.array([[128]], np.uint8), np.array([[128]], np.uint8))[1] == False).all()  # gt > 128


def test_quantisation_truncates():
    q = M.quantise_like_reference(np.array([0.0, 10.0, -10.0], np.float32))
    assert q.tolist() == [127, 254, 0]


def test_weighted_f_tolerates_boundary_error_more_than_far_error():
    gt = _square()
    near = gt.copy()
    near[15, 16:48] = 255  # false positives hugging the object
    far = gt.copy()
    far[2, 16:48] = 255    # the same number of false positives far away
    assert M.score_pair(near, gt)["wfm"] > M.score_pair(far, gt)["wfm"]


def test_aggregate_is_plain_mean():
    rows = [{"sm": 1.0, "wfm": 0.5, "mae": 0.1, "em": 0.9, "fm": 0.7}, {"sm": 0.0, "wfm": 0.5, "mae": 0.3, "em": 0.7, "fm": 0.1}]
    agg = M.aggregate(rows)
    assert agg == pytest.approx({"s_alpha": 0.5, "weighted_f": 0.5, "mae": 0.2, "e_phi": 0.8, "mean_f": 0.4})


def _feature_transform_int(fg):
    """Host restatement of the integer algorithm of csrc/metrics.cu (ft_columns_kernel + ft_rows_kernel): nearest
    foreground pixel per pixel, with scipy's choice among equidistant candidates."""
    H, W = fg.shape
    fy = -np.ones((H, W), np.int64)
    for x in range(W):
        last = -1
        for y in range(H):
            if fg[y, x]:
                last = y
            fy[y, x] = last
        nxt = -1
        for y in range(H - 1, -1, -1):
            if fg[y, x]:
                nxt = y
            up = fy[y, x]
            if nxt >= 0 and (up < 0 or nxt - y < y - up):
                fy[y, x] = nxt
    out = -np.ones((H, W), np.int64)
    for y in range(H):
        g = []
        for ii in range(W):
            if fy[y, ii] < 0:
                continue
            wR = (fy[y, ii] - y) ** 2
            while len(g) >= 2:
                i1, i2 = g[-1], g[-2]
                a, b = i1 - i2, ii - i1
                c = a + b
                uR, vR = (fy[y, i2] - y) ** 2, (fy[y, i1] - y) ** 2
                if c * vR - b * uR - a * wR - a * b * c <= 0:
                    break
                g.pop()
            g.append(ii)
        if not g:
            continue
        l = 0
        for ii in range(W):
            d1 = (g[l] - ii) ** 2 + (fy[y, g[l]] - y) ** 2
            while l < len(g) - 1:
                d2 = (g[l + 1] - ii) ** 2 + (fy[y, g[l + 1]] - y) ** 2
                if d1 <= d2:
                    break
                d1 = d2
                l += 1
            out[y, ii] = fy[y, g[l]] * W + g[l]
    return out


def test_integer_feature_transform_reproduces_scipy_including_ties():
    """The GPU weighted-F pass reads the prediction error AT the nearest foreground pixel, so the choice among
    equidistant pixels matters; the kernels' integer scan must make scipy's choice (pinned here on the host, and
    bit-for-bit on the GPU in tests/test_gpu_metrics.py)."""
    from scipy.ndimage import distance_transform_edt

    rng = np.random.default_rng(0)
    for trial in range(40):
        H, W = (int(v) for v in rng.integers(5, 36, 2))
        kind = trial % 4
        if kind == 0:
            fg = rng.random((H, W)) < 0.05
        elif kind == 1:
            fg = rng.random((H, W)) < 0.5
        elif kind == 2:
            yy, xx = np.mgrid[:H, :W]
            fg = ((yy - H / 2) ** 2 / (H / 3) ** 2 + (xx - W / 2) ** 2 / (W / 4) ** 2) < 1
        else:
            fg = np.zeros((H, W), bool)
            fg[rng.integers(0, H), rng.integers(0, W)] = True
            fg[rng.integers(0, H), rng.integers(0, W)] = True
        if not fg.any():
            continue
        _, idx = distance_transform_edt(~fg, return_indices=True)
        assert np.array_equal(_feature_transform_int(fg), idx[0] * W + idx[1]), trial
