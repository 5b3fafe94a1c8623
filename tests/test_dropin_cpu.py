"""CPU: the drop-in surface the reference's callers rely on (config access, refusals, checkpoint layouts)."""
import os

import numpy as np
import pytest
import torch

from diffusion_pruning_b200 import UNet2DConditionModelGated, UNet2DConditionModelPruned
from diffusion_pruning_b200.synthetic import synthetic_codes
from oracle.unet_oracle import GatedUNetOracle, UNetConfig, seeded_init

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)


def test_config_is_attribute_and_item_mapping_and_read_only():
    m = UNet2DConditionModelGated(**TINY, sample_size=32, upcast_attention=True, addition_embed_type_num_heads=64)
    assert m.config.sample_size == 32 and m.config["in_channels"] == 4 and m.config.cross_attention_dim == 128
    assert m.config.upcast_attention is True            # inert key preserved, not swallowed
    assert m.dtype == torch.float32 and m.device.type == "cpu"
    with pytest.raises(AttributeError):
        m.config.in_channels = 8
    m.register_to_config(encoder_hid_dim_type=None)
    m.enable_gradient_checkpointing()
    assert m.gradient_checkpointing
    with pytest.raises(NotImplementedError):
        m.enable_xformers_memory_efficient_attention()


@pytest.mark.parametrize("key,val", [("dual_cross_attention", True), ("resnet_time_scale_shift", "scale_shift"),
                                     ("class_embed_type", "timestep"), ("use_linear_projection", False),
                                     ("addition_embed_type", "text_time")])
def test_unsupported_config_values_raise_instead_of_being_swallowed(key, val):
    with pytest.raises(NotImplementedError):
        UNet2DConditionModelGated(**TINY, **{key: val})


def test_calc_macs_needs_a_forward_or_count_macs():
    m = UNet2DConditionModelGated(**TINY)
    with pytest.raises(RuntimeError, match="forward"):
        m.calc_macs()
    m.count_macs(16, 16)
    assert m.calc_macs()["total_macs"] > 0


def _pruned_pair():
    o = GatedUNetOracle(UNetConfig.tiny()).eval()
    seeded_init(o, 0, 0.1)
    m = UNet2DConditionModelPruned(**TINY)
    m.load_state_dict(o.state_dict())
    code = synthetic_codes(m.get_structure(), 8)[3:4]
    return o, m, code


def test_sliced_layout_has_the_reference_pruned_models_keys_and_shapes():
    """Keys / shapes recorded from the reference's own prune sweep (unet_2d_conditional.py:2425-2436, golden fixture)."""
    _, m, code = _pruned_pair()
    m.prune_to(code * 0.9 + 0.05)
    sl = m.sliced_state_dict()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "unet_ref.npz"))
    keys, shapes = list(gold["pruned_state_keys"]), gold["pruned_state_shapes"]
    assert set(sl) == set(keys)
    for k, sh in zip(keys, shapes):
        assert tuple(int(x) for x in sh[:sl[k].dim()]) == tuple(sl[k].shape), k


def test_sliced_values_equal_the_physically_pruned_oracle():
    import copy
    from diffusion_pruning_b200.synthetic import split_arch
    o, m, code = _pruned_pair()
    m.prune_to(code * 0.9 + 0.05)
    sl = m.sliced_state_dict()
    p = copy.deepcopy(o)
    p.set_structure(split_arch(code.clone(), p.get_structure()))
    p.prune()                                       # slices the ResNets physically (blocks.py:424-465, :641-697)
    psd = p.state_dict()
    n = 0
    for k, v in sl.items():
        if ".resnets." in k and k in psd and psd[k].shape == v.shape:
            assert torch.equal(psd[k], v), k
            n += 1
    assert n > 100


def test_sliced_checkpoint_round_trips_through_from_pretrained(tmp_path):
    _, m, code = _pruned_pair()
    arch = code * 0.9 + 0.05
    m.prune_to(arch)
    d = tmp_path / "ckpt"
    m.save_pretrained(os.path.join(d, "unet"), sliced=True)
    torch.save(arch, os.path.join(d, "arch_vector.pt"))       # where FineTuner puts it (trainer.py:1452, :1659-1661)
    m2 = UNet2DConditionModelPruned.from_pretrained(os.fspath(d), subfolder="unet")
    a, b = m.sliced_state_dict(), m2.sliced_state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    # pruned rows / columns are zero-filled in the dense parameters, kept ones are in place
    dense = m2.state_dict()
    k = "down_blocks.0.resnets.0.conv1.weight"
    assert dense[k].shape == m.state_dict()[k].shape
    # without a code a sliced file must fail loudly
    m3 = UNet2DConditionModelPruned(**TINY)
    with pytest.raises(RuntimeError, match="sliced"):
        m3.load_state_dict(a)
    # dense checkpoints still load
    m.save_pretrained(os.path.join(d, "unet_dense"))
    m4 = UNet2DConditionModelPruned.from_pretrained(os.path.join(d, "unet_dense"), arch_vector=arch)
    assert torch.equal(m4.state_dict()[k], m.state_dict()[k])
