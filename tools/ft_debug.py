#!/usr/bin/env python
"""Timing breakdown of the fine-tune step's host-side phases on the full-size model (small batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_pruning_b200 import finetune as FT, pruning_step as PS
from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
from diffusion_pruning_b200.unet import UNet2DConditionModelGated

dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet2DConditionModelGated(); teacher = UNet2DConditionModelGated()
teacher.load_state_dict(unet.state_dict()); teacher.eval(); teacher.freeze(); teacher.set_all_ones_structure(1, device=dev)
unet.enable_weight_training(True)
st = unet.get_structure()
code = synthetic_codes(st, 8)[3:4].float().to(dev)
B = int(os.environ.get("B", "4"))
g = torch.Generator().manual_seed(1)
batch = {"noisy_latents": torch.randn(B, 4, 64, 64, generator=g).to(dev), "timesteps": torch.randint(0, 1000, (B,), generator=g).to(dev),
         "target": torch.randn(B, 4, 64, 64, generator=g).to(dev), "encoder_hidden_states": torch.randn(B, 77, 1024, generator=g).to(dev)}
cfg = FT.FinetuneLossConfig()
taps, ttaps = PS.BlockTaps(unet), PS.BlockTaps(teacher)
opt = torch.optim.AdamW([p for p in unet.parameters()], lr=1e-5, weight_decay=0.0, fused=True)
acp = PS.alphas_cumprod().to(dev)


def sync():
    torch.cuda.synchronize(); return time.perf_counter()

for it in range(6):
    t0 = sync()
    unet.set_structure(split_arch(code.clone(), st))
    eng = unet._get_train_engine(dev)
    t1 = sync()
    if getattr(eng, "_ft_packs_ready", False):
        eng.refresh_packs()  # what the forward will do again (measured separately here)
    t2 = sync()
    out = FT.finetune_step(unet, teacher, batch, cfg, taps, ttaps, acp=acp)
    t3 = sync()
    opt.zero_grad(set_to_none=True)
    out["loss"].backward()
    t4 = sync()
    opt.step()
    t5 = sync()
    print(f"it{it}: set_structure {1e3*(t1-t0):.1f}  refresh {1e3*(t2-t1):.1f}  fwd(teacher+student+losses) {1e3*(t3-t2):.1f}  "
          f"bwd {1e3*(t4-t3):.1f}  adamw {1e3*(t5-t4):.1f}  packs {len(eng._packs)} graph_n {eng._pack_graph_n}", flush=True)
