"""Pins the restated Hiera trunk (oracle/hiera.py) against the independent HF port that ships in the
image (transformers.models.sam2): same weights under a key remap -> same 4 feature maps, fp32."""
import pytest
import torch

from oracle.hiera import HieraConfig, HieraTrunk, block_specs, hf_key_to_upstream


def _hf_model(cfg: HieraConfig):
    from transformers.models.sam2.configuration_sam2 import Sam2HieraDetConfig
    from transformers.models.sam2.modeling_sam2 import Sam2HieraDetModel

    hfc = Sam2HieraDetConfig(
        hidden_size=cfg.embed_dim, blocks_per_stage=list(cfg.stages), embed_dim_per_stage=cfg.stage_dims,
        num_attention_heads_per_stage=cfg.stage_heads, window_size_per_stage=list(cfg.window_spec),
        global_attention_blocks=list(cfg.global_att_blocks),
        window_positional_embedding_background_size=list(cfg.window_pos_embed_bkg_spatial_size))
    return Sam2HieraDetModel(hfc).eval()


@pytest.mark.parametrize("cfg,size", [
    (HieraConfig(), 256),                                                   # the real Hiera-L geometry
    (HieraConfig(stages=(1, 2, 3, 2), global_att_blocks=(4,)), 512),        # shallow, default resolution
    # 352: the stage-3 / stage-4 token grids (22 / 11) do not tile into 16 / 8 windows -> padded windows (HF:395-399),
    # incl. the query-pooling block at the stage change (window 16 on the 22 grid, pooled to 8 on the 11 grid)
    (HieraConfig(stages=(1, 2, 4, 2), global_att_blocks=(5,)), 352),
    (HieraConfig(stages=(1, 2, 3, 2), global_att_blocks=(4,)), 384),
])
def test_trunk_matches_hf_port(cfg, size):
    torch.manual_seed(0)
    hf = _hf_model(cfg)
    with torch.no_grad():
        for n, p in hf.named_parameters():
            p.normal_(0, 0.05 if p.dim() > 1 else 0.1)
            if "layer_norm" in n and n.endswith("weight"):
                p.add_(1.0)
    trunk = HieraTrunk(cfg)
    trunk.load_state_dict({hf_key_to_upstream(k): v for k, v in hf.state_dict().items()}, strict=True)
    x = torch.randn(1, 3, size, size)
    with torch.no_grad():
        ours = trunk(x)
        theirs = hf(pixel_values=x).intermediate_hidden_states
    assert len(ours) == 4
    for o, t in zip(ours, theirs):
        t = t.permute(0, 3, 1, 2)
        assert o.shape == t.shape
        assert (o - t).abs().max() <= 1e-4 * max(1.0, float(t.abs().max()))


def test_block_geometry():
    specs = block_specs(HieraConfig())
    assert len(specs) == 48
    assert [s.index for s in specs if s.q_stride] == [2, 8, 44]
    assert [s.index for s in specs if s.window == 0] == [23, 33, 43]
    # the window size lags the stage change by one block
    assert (specs[2].window, specs[3].window) == (8, 4)
    assert (specs[8].window, specs[9].window) == (4, 16)
    assert (specs[44].window, specs[45].window) == (16, 8)
    assert all(s.dim_out // s.heads == 72 for s in specs)
    assert sum(p.numel() for p in HieraTrunk().parameters()) == 212_149_296


def test_input_validation():
    from oracle.spegnet import spegnet_forward

    with pytest.raises(ValueError):
        spegnet_forward({}, torch.zeros(3, 64, 64))
    with pytest.raises(ValueError):
        spegnet_forward({}, torch.zeros(1, 3, 100, 100))
