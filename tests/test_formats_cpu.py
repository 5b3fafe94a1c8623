"""On-disk formats of the prune -> finetune -> generate workflow (host logic only) and the oracle's
restatement of prune() (blocks.py:424-465). No GPU needed: nothing here runs a forward of the product."""
import copy
import json
import os

import torch

from diffusion_pruning_b200 import (HyperStructure, StructureVectorQuantizer, UNet2DConditionModelGated,
                                    UNet2DConditionModelPruned)
from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
from oracle.unet_oracle import GatedUNetOracle, UNetConfig, seeded_init

TINY = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)


def _tiny_oracle(beta_std):
    o = GatedUNetOracle(UNetConfig.tiny()).eval()
    seeded_init(o, 0, beta_std)
    return o


def test_unet_diffusers_layout_roundtrip(tmp_path):
    o = _tiny_oracle(0.1)
    m = UNet2DConditionModelGated(**TINY)
    m.load_state_dict(o.state_dict())
    m.save_pretrained(os.path.join(tmp_path, "unet"))
    assert sorted(os.listdir(os.path.join(tmp_path, "unet"))) == ["config.json", "diffusion_pytorch_model.safetensors"]
    m2 = UNet2DConditionModelGated.from_pretrained(str(tmp_path), subfolder="unet")
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd.keys()) == list(sd2.keys()) == list(o.state_dict().keys())
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert m2.config["block_out_channels"] == (64, 128, 256, 256)


def test_stock_sd_config_is_mapped_to_gated_blocks(tmp_path):
    m = UNet2DConditionModelGated(**TINY)
    d = os.path.join(tmp_path, "unet")
    m.save_pretrained(d)
    cfg = json.load(open(os.path.join(d, "config.json")))
    cfg.update(_class_name="UNet2DConditionModel", _diffusers_version="0.23.1", act_fn="silu",
               down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], mid_block_type="UNetMidBlock2DCrossAttn",
               up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, use_linear_projection=True)
    cfg.pop("gated_ff")
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    m2 = UNet2DConditionModelGated.from_pretrained(d, ff_gate_width=32)
    assert m2.config["down_block_types"][0] == "CrossAttnDownBlock2DHalfGated"
    assert m2.get_structure() == m.get_structure()


def test_pruned_expert_loads_arch_vector_from_checkpoint(tmp_path):
    m = UNet2DConditionModelGated(**TINY)
    m.save_pretrained(os.path.join(tmp_path, "unet"))
    st = m.get_structure()
    codes = synthetic_codes(st, 8, seed=2).float()
    soft = codes[5:6] * 0.93 + 0.02  # quantizer_embeddings.pt rows are soft gumbel-sigmoid values (trainer.py:274)
    torch.save(soft, os.path.join(tmp_path, "arch_vector.pt"))
    p = UNet2DConditionModelPruned.from_pretrained(str(tmp_path), subfolder="unet")
    assert p.pruned_semantics and torch.equal(p.arch_vector, soft)
    flat_w, flat_d = p._flat_gates
    got = torch.cat([w.reshape(1, -1) for w in flat_w] + [d.reshape(1, 1) for d in flat_d], dim=1)
    assert torch.equal(got, codes[5:6])  # thresholded at 0.5, exact 0/1 (hard_concrete in prune())
    p2 = UNet2DConditionModelPruned.from_pretrained(str(tmp_path), subfolder="unet", arch_vector=codes[1:2])
    assert torch.equal(p2.arch_vector, codes[1:2])
    assert list(p.state_dict().keys()) == list(m.state_dict().keys())  # dense diffusers keys, no gate tensors


def test_router_checkpoint_files_roundtrip(tmp_path):
    st = UNet2DConditionModelGated(**TINY).get_structure()
    hyper = HyperStructure(structure=st, input_dim=32, wn_flag=False, linear_bias=True)
    quant = StructureVectorQuantizer(n_e=4, structure=st, beta=0.25, temperature=0.4, base=3, depth_order=list(range(14)),
                                     non_zero_width=True, resource_aware_normalization=False, optimal_transport=True)
    hyper.save_pretrained(os.path.join(tmp_path, "hypernet"))
    quant.save_pretrained(os.path.join(tmp_path, "quantizer"))
    torch.save(quant.embedding_gs, os.path.join(tmp_path, "quantizer_embeddings.pt"))  # trainer.py:274
    h2 = HyperStructure.from_pretrained(str(tmp_path), subfolder="hypernet")
    q2 = StructureVectorQuantizer.from_pretrained(str(tmp_path), subfolder="quantizer")
    assert all(torch.equal(a, b) for a, b in zip(hyper.state_dict().values(), h2.state_dict().values()))
    assert all(torch.equal(a, b) for a, b in zip(quant.state_dict().values(), q2.state_dict().values()))
    emb = torch.load(os.path.join(tmp_path, "quantizer_embeddings.pt"), map_location="cpu")
    arch_v = emb[2 % emb.shape[0]].unsqueeze(0)  # trainer.py:1447-1448
    assert arch_v.shape == (1, quant.vq_embed_dim)


def test_oracle_prune_equals_gated_iff_groupnorm_beta_is_zero():
    """prune() (blocks.py:451-463) deletes the channels of gated-off GroupNorm groups, the gate
    leaves silu(beta) flowing into conv2. The two agree exactly when beta = 0 and differ otherwise."""
    g = torch.Generator().manual_seed(1)
    x, c, t = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, 128, generator=g), torch.tensor([981, 21])
    for beta_std, same in ((0.0, True), (0.1, False)):
        o = _tiny_oracle(beta_std)
        st = o.get_structure()
        code = synthetic_codes(st, 8, seed=2)[3:4].float()
        o.set_structure(split_arch(code.clone(), st))
        with torch.no_grad():
            y_gated = o(x, t, c)
        p = copy.deepcopy(o)
        p.prune()
        n_sliced = sum(int(getattr(m, "pruned", False)) for m in p.modules())
        assert n_sliced > 0
        with torch.no_grad():
            y_pruned = p(x, t, c)
        diff = (y_gated - y_pruned).abs().max().item()
        assert (diff < 1e-4) if same else (diff > 1e-3), (beta_std, diff)


def test_macs_device_tables_cover_the_gate_matrix_and_reproduce_the_closed_form():
    """Host logic of the K7 kernel (macs._device_tables): the aptp_macs_gate / aptp_macs_sub records must tile the
    [B, 1620-like] gate matrix exactly once (width columns, then depth columns, get_structure order) and carry the same
    constants as the torch closed form (evaluated here on CPU gates)."""
    import numpy as np
    from diffusion_pruning_b200 import macs as M
    m = UNet2DConditionModelGated(**TINY)
    m.count_macs(16, 16)
    tab = M._device_tables(m, torch.device("cpu"))
    from diffusion_pruning_b200 import kernels as K
    gates = tab["gates"].numpy().view(K.MACS_GATE_DTYPE)
    subs = tab["subs"].numpy().view(K.MACS_SUB_DTYPE)
    st = m.get_structure()
    widths = [w for ws in st["width"] for w in ws]
    n_depth = sum(1 for d in st["depth"] if d == [1])
    assert len(gates) == len(widths) == tab["n_gates"] and len(subs) == len(st["width"]) == tab["n_subs"]
    assert gates["width"].tolist() == widths
    assert gates["col"].tolist() == np.concatenate([[0], np.cumsum(widths)[:-1]]).tolist()
    dcols = [int(c) for c in subs["depth_col"] if c >= 0]
    assert dcols == list(range(sum(widths), sum(widths) + n_depth)) and tab["dim"] == sum(widths) + n_depth
    assert subs["first_gate"].tolist() == np.concatenate([[0], np.cumsum(subs["n_gates"])[:-1]]).tolist()
    # all-ones gates: cur_prunable = sum of gate MACs + fixed parts of the depth-gated sub-blocks; totals match the closed form
    d = M._calc_macs_torch(m)
    depth_fixed = float(sum(s["fixed"] for s in subs if s["depth_col"] >= 0))
    assert abs(float(gates["macs"].sum()) + depth_fixed - float(d["cur_prunable_macs"])) <= 1e-6 * float(d["cur_prunable_macs"])
    assert abs(tab["total"] - d["total_macs"]) <= 1e-9 * d["total_macs"]
    assert abs(tab["prunable"] - d["prunable_macs"]) <= 1e-9 * d["prunable_macs"]
