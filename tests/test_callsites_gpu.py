"""GPU: the reference's own call sequences (trainer.py:1257-1296, pruning_pipelines.py:746-824) run UNCHANGED against the
drop-in classes, and their results agree with the same sequences run on the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_count_macs_sequence_and_pipeline_loop_match_oracle():
    import callsite_checks as C
    from diffusion_pruning_b200.synthetic import split_arch
    from oracle.unet_oracle import GatedUNetOracle, UNetConfig
    dev = torch.device("cuda")
    unet, hyper_net, quantizer = C.build(dev)
    macs = C.replay_count_macs(unet, hyper_net, quantizer, dev)
    # oracle side of the MAC numbers (itself pinned to the reference's op_counter run, tests/test_oracle_unet_pinned.py)
    oracle = GatedUNetOracle(UNetConfig.tiny()).eval()
    oracle.load_state_dict({k: v.float().cpu() for k, v in unet.state_dict().items()})
    oracle.count_macs(16, 16, 77)
    oracle.set_all_ones(1)
    ref = oracle.calc_macs()
    assert macs["total_macs"] == float(ref["total_macs"]) and macs["prunable_macs"] == float(ref["prunable_macs"])
    assert abs(macs["cur_prunable_macs"] - float(ref["cur_prunable_macs"])) <= 1e-6 * macs["cur_prunable_macs"]

    latents, idx, ratios, zq, trace = C.replay_pipeline(unet, hyper_net, quantizer, dev)
    assert latents.shape == (3, 4, 16, 16) and torch.isfinite(latents).all()
    assert ratios.shape[0] == 3 and ((ratios > 0) & (ratios <= 1.0 + 1e-6)).all()
    # first U-Net call of the loop vs the oracle on the same routed structure and inputs
    g = torch.Generator().manual_seed(5)
    cond = torch.randn(3, 77, 128, generator=g)
    neg = torch.randn(1, 77, 128, generator=g).expand(3, -1, -1)
    _ = torch.randn(3, 32, generator=g)
    lat0 = torch.randn(3, 4, 16, 16, generator=g)
    st = oracle.get_structure()
    oracle.set_structure(split_arch(zq.detach().float().cpu(), st))
    t0 = torch.tensor([(1000 // 3) * 2 + 1])
    with torch.no_grad():
        ref_pred = oracle(torch.cat([lat0] * 2), t0, torch.cat([neg, cond]))
    got = trace[0].float().cpu()                     # raw [uncond; cond] prediction of the first loop iteration
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref_pred.flatten(), dim=0).item()
    err = (got - ref_pred).abs().max().item() / max(1.0, ref_pred.abs().max().item())
    assert cos >= 0.9999 and err <= 2e-2, (cos, err)
    # after guidance 7.5 the bf16 difference of the two halves is amplified 7.5x: looser, but the same direction
    gu, gc = got.chunk(2)
    ru, rc = ref_pred.chunk(2)
    cosg = torch.nn.functional.cosine_similarity((gu + 7.5 * (gc - gu)).flatten(), (ru + 7.5 * (rc - ru)).flatten(), dim=0)
    assert cosg.item() >= 0.995, cosg.item()


def test_api_surface_the_callers_touch():
    import callsite_checks as C
    dev = torch.device("cuda")
    unet, _, _ = C.build(dev)
    assert unet.config.sample_size == 16 and unet.config["in_channels"] == 4       # pruning_pipelines.py:708, :772
    assert unet.dtype == torch.float32 and unet.device.type == "cuda"
    unet.enable_gradient_checkpointing()                                            # trainer.py:160
    with pytest.raises(NotImplementedError, match="head gating"):
        unet.enable_xformers_memory_efficient_attention()                           # trainer.py:154 must not go silent
    unet.register_to_config(encoder_hid_dim_type=None)
    with pytest.raises(AttributeError):
        unet.config.sample_size = 3
