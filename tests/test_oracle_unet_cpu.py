"""Structural anchors of the U-Net oracle (oracle/unet_oracle.py is a restatement: diffusers 0.23.1 is not installable
here). What CAN be pinned without it: the public SD-2.1 U-Net parameter count and state-dict layout, the
reference's own gate layout, and known-answer / self-consistency properties of the restated arithmetic."""
import torch

from diffusion_pruning_b200 import UNet2DConditionModelGated
from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
from oracle.unet_oracle import GatedUNetOracle, UNetConfig, seeded_init, timestep_sinusoid

SD21_UNET_PARAMS = 865_910_724   # stabilityai/stable-diffusion-2-1, unet/diffusion_pytorch_model.safetensors
SD21_UNET_TENSORS = 686


def _full_on_meta(cls):
    with torch.device("meta"):
        return cls()


def test_full_size_oracle_has_the_public_sd21_parameter_count_and_keys():
    o = _full_on_meta(GatedUNetOracle)
    sd = o.state_dict()
    assert sum(v.numel() for v in sd.values()) == SD21_UNET_PARAMS
    assert len(sd) == SD21_UNET_TENSORS
    expect = {  # diffusers key -> shape (use_linear_projection=True: proj_in / proj_out are 2-D)
        "conv_in.weight": (320, 4, 3, 3),
        "time_embedding.linear_1.weight": (1280, 320),
        "time_embedding.linear_2.bias": (1280,),
        "down_blocks.0.resnets.0.time_emb_proj.weight": (320, 1280),
        "down_blocks.0.attentions.0.norm.weight": (320,),
        "down_blocks.0.attentions.0.proj_in.weight": (320, 320),
        "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight": (320, 320),
        "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight": (320, 1024),
        "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_out.0.bias": (320,),
        "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight": (2560, 320),
        "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.2.weight": (320, 1280),
        "down_blocks.0.downsamplers.0.conv.weight": (320, 320, 3, 3),
        "down_blocks.1.resnets.0.conv_shortcut.weight": (640, 320, 1, 1),
        "down_blocks.3.resnets.1.conv2.weight": (1280, 1280, 3, 3),
        "mid_block.attentions.0.transformer_blocks.0.norm3.weight": (1280,),
        "mid_block.resnets.1.norm2.bias": (1280,),
        "up_blocks.0.resnets.2.conv_shortcut.weight": (1280, 2560, 1, 1),
        "up_blocks.0.upsamplers.0.conv.weight": (1280, 1280, 3, 3),
        "up_blocks.3.resnets.0.conv1.weight": (320, 960, 3, 3),
        "up_blocks.3.attentions.2.proj_out.bias": (320,),
        "conv_norm_out.weight": (320,),
        "conv_out.weight": (4, 320, 3, 3),
    }
    for k, shp in expect.items():
        assert k in sd and tuple(sd[k].shape) == shp, (k, tuple(sd[k].shape) if k in sd else None)
    assert not any("gate" in k for k in sd), "gates add no parameters / buffers (gates.py:13)"


def test_product_module_tree_has_the_same_state_dict_layout():
    o, m = _full_on_meta(GatedUNetOracle), _full_on_meta(UNet2DConditionModelGated)
    so, sm = o.state_dict(), m.state_dict()
    assert list(so.keys()) == list(sm.keys())
    assert all(so[k].shape == sm[k].shape for k in so)


def test_gate_layout_totals_match_the_reference_structure():
    m = _full_on_meta(UNet2DConditionModelGated)
    st = m.get_structure()
    widths = [w for ws in st["width"] for w in ws]
    assert len(widths) == 70 and sum(widths) == 1606
    assert sum(1 for d in st["depth"] if d == [1]) == 14 and len(st["width"]) == 38
    assert _full_on_meta(GatedUNetOracle).get_structure() == st


def test_timestep_embedding_known_answers():
    e = timestep_sinusoid(torch.tensor([0, 1000]), 320)
    assert torch.equal(e[0], torch.cat([torch.ones(160), torch.zeros(160)]))      # flip_sin_to_cos: cos half first
    assert abs(e[1, 160].item() - torch.sin(torch.tensor(1000.0)).item()) < 1e-6  # frequency 10000^0 = 1
    assert abs(e[1, 319].item() - torch.sin(torch.tensor(1000.0 * 10000 ** (-159 / 160))).item()) < 1e-5


def test_cfg_doubled_batch_equals_two_halves_and_hard_gates_only_remove_work():
    o = GatedUNetOracle(UNetConfig.tiny()).eval()
    seeded_init(o, 0, 0.1)
    st = o.get_structure()
    codes = synthetic_codes(st, 8).float()
    g = torch.Generator().manual_seed(3)
    x, c, t = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, 128, generator=g), torch.tensor([981, 21])
    with torch.no_grad():
        o.set_structure(split_arch(codes[[1, 5]].clone(), st))
        y = o(x, t, c)
        # gates for 2 prompts tiled over a batch of 4 = [uncond; cond] (gates.py:18-19)
        y4 = o(torch.cat([x, x]), torch.cat([t, t]), torch.cat([c, c]))
        assert torch.allclose(y4[:2], y, atol=1e-5) and torch.allclose(y4[2:], y, atol=1e-5)
        # a sample's output depends on its own code only: same sample, other batch neighbours
        o.set_structure(split_arch(codes[[1, 2]].clone(), st))
        y2 = o(x, t, c)
        assert torch.allclose(y2[0], y[0], atol=1e-5) and not torch.allclose(y2[1], y[1], atol=1e-3)
        # all-ones gates: depth lerp and width gates are exact identities on the dense path
        o.set_all_ones(2)
        y_ones = o(x, t, c)
        o.set_all_ones(1)
        assert torch.allclose(o(x, t, c), y_ones, atol=1e-6)
