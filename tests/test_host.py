"""CPU-side checks (no GPU): the C-ABI libraries load and export every symbol the header declares, the
drop-in module mirrors the reference's constructor / state-dict / error contract, and the product never
falls back to a CPU path."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = {"encoder": {"config_path": "configs/sam2.1/sam2.1_hiera_l.yaml",
                   "checkpoint_path": "./checkpoints/sam2.1_hiera_large.pt", "variant": "large"},
       "name": "spegnet", "image_processing": {"target_size": 512}}


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spegnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spg_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("variant", ["fp16", "bf16"])
def test_library_exports_every_declared_symbol(variant):
    from spegnet_b200 import _lib

    path = _lib.LIB_PATHS[variant]
    assert os.path.exists(path), "run `python -m spegnet_b200.build` (or __graft_entry__.build()) first"
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/spegnet_b200.h but not exported by {path}"
    # and the Python binding covers the same set
    assert sorted(_lib.EXPORTS) == declared
    loaded = _lib.load(variant)
    assert loaded.spg_version() >= 100
    assert bool(loaded.spg_half_is_fp16()) == (variant == "fp16")
    assert loaded.spg_last_error() == b"" or isinstance(loaded.spg_last_error(), bytes)


def test_argument_errors_surface_without_a_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the build box."""
    from spegnet_b200 import _lib

    lib = _lib.load("fp16")
    ep = _lib.Epilogue()
    rc = lib.spg_linear_h16(None, None, 128, 64, 64, ctypes.byref(ep), None)
    assert rc == -1 and b"NULL" in lib.spg_last_error()
    rc = lib.spg_linear_h16(ctypes.c_void_p(16), ctypes.c_void_p(16), 128, 60, 64, ctypes.byref(ep), None)
    assert rc == -1 and b"multiple of 16" in lib.spg_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "spg_linear_h16", "fp16")


@pytest.fixture(scope="module")
def model():
    from spegnet_b200 import SPEGNet

    return SPEGNet(CFG)


def test_state_dict_schema_matches_reference(model, spread_sd):
    """Key names and shapes are the reference's (models/spegnet.py:94-135 + sam2 trunk names), so a
    reference checkpoint's ['model_state_dict'] loads with strict=True."""
    ours = model.state_dict()
    assert set(ours) == set(spread_sd)
    for k, v in spread_sd.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    assert sum(p.numel() for p in model.parameters()) == 215_442_100
    res = model.load_state_dict(spread_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # trainer-style parameter grouping by name still works (engine/trainer.py:284-294)
    names = [n for n, _ in model.named_parameters()]
    assert any("encoder" in n for n in names) and any("bn" in n for n in names) and any("norm" in n for n in names)


def test_constructor_and_input_errors_match_reference(model):
    from spegnet_b200 import SPEGNet

    with pytest.raises(ValueError):  # models/feature_encoding.py:150-151
        SPEGNet({"encoder": {"config_path": "", "checkpoint_path": "", "variant": "gigantic"}})
    with pytest.raises(KeyError):
        SPEGNet({})
    with pytest.raises(ValueError, match="4D"):  # models/feature_encoding.py:230-231
        model(torch.zeros(3, 512, 512))
    with pytest.raises(ValueError, match="divisible by 32"):  # :232-233
        model(torch.zeros(1, 3, 500, 500))
    with pytest.raises(NotImplementedError):
        model.train()


def test_no_cpu_fallback(model):
    """A CPU tensor must fail loudly instead of silently running an eager / oracle path."""
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 512, 512))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spegnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_trunk_geometry_agrees_with_oracle():
    from oracle.hiera import HieraConfig, block_specs
    from spegnet_b200.schema import trunk_blocks

    for a, b in zip(trunk_blocks(), block_specs(HieraConfig())):
        assert (a.index, a.stage, a.dim_in, a.dim_out, a.heads, a.window, a.q_pool) == (
            b.index, b.stage, b.dim_in, b.dim_out, b.heads, b.window, bool(b.q_stride))


def test_up2_phase_weights_reproduce_interpolate_then_conv():
    """Host-side weight folding of the fused `bilinear x2 -> conv3x3` kernel (spg_conv3x3_up2_h16): a low-resolution
    zero-padded conv with one weight set per row class, plus the border-column correction, equals
    conv2d(interpolate(x, 2x, bilinear), w, padding=1) (models/object_detection.py:219,230) to fp64 rounding."""
    import torch.nn.functional as F

    from spegnet_b200.model import up2_operators, up2_phase_weights

    R = up2_operators()
    assert float((R[0] - R[1])[:, :, 2].abs().max()) == 0.0 and float((R[2] - R[1])[:, :, 0].abs().max()) == 0.0
    g = torch.Generator().manual_seed(0)
    B, H, W, Ci, Co = 2, 6, 7, 4, 3
    x = torch.randn(B, Ci, H, W, dtype=torch.float64, generator=g)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), w, padding=1)
    main, dwl, dwr = up2_phase_weights(w)
    main = main.view(3, 4 * Co, 9 * Ci)
    xp = F.pad(x, (1, 1, 1, 1))
    cols = torch.stack([xp[:, :, dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)], dim=1)
    A = cols.permute(0, 3, 4, 1, 2).reshape(B, H, W, 9 * Ci)
    out = torch.zeros(B, H, W, 4 * Co, dtype=torch.float64)
    cls_of = lambda y: 0 if y == 0 else (2 if y == H - 1 else 1)  # noqa: E731
    for y in range(H):
        out[:, y] = A[:, y] @ main[cls_of(y)].t()
    for col, dW in ((0, dwl), (W - 1, dwr)):
        for y in range(H):
            a = torch.zeros(B, 9 * Ci, dtype=torch.float64)
            for dy in range(3):
                if 0 <= y + dy - 1 < H:
                    k = (cls_of(y) * 3 + dy) * Ci
                    a[:, k:k + Ci] = x[:, :, y + dy - 1, col]
            out[:, y, col] += a @ dW.t()
    got = out.reshape(B, H, W, 2, 2, Co).permute(0, 5, 1, 3, 2, 4).reshape(B, Co, 2 * H, 2 * W)
    assert float((got - ref).abs().max()) < 1e-12


def test_collate_fn_mirrors_the_reference():
    """utils/data_loader.py:177-212: stacked images, ragged masks kept as a list, 'names' for test samples / 'edges' for
    training samples, ValueError on an empty batch."""
    import torch

    from spegnet_b200.batching import collate_fn

    items = [{"image": torch.zeros(3, 8, 8) + i, "mask": torch.ones(1, 5 + i, 7), "name": f"img{i}.png"} for i in range(3)]
    out = collate_fn(items)
    assert tuple(out["images"].shape) == (3, 3, 8, 8) and float(out["images"][2].mean()) == 2.0
    assert [tuple(m.shape) for m in out["masks"]] == [(1, 5, 7), (1, 6, 7), (1, 7, 7)]
    assert out["names"] == ["img0.png", "img1.png", "img2.png"] and "edges" not in out
    train = collate_fn([dict(it, edge=torch.zeros(1, 4, 4)) for it in items])
    assert len(train["edges"]) == 3 and "names" not in train
    with pytest.raises(ValueError, match="Empty batch"):
        collate_fn([])


def test_tools_and_entry_points_compile():
    """Every development tool, the bench and the entry module are at least syntactically valid on this interpreter
    (they only run on the GPU box, where a typo would cost a whole call)."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"),
                                                                       os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 10
    for f in files:
        with open(f) as fh:
            compile(fh.read(), f, "exec")
