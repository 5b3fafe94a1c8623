"""torch.profiler view of one pruning train step: which PyTorch (non-aptp) ops still run on the GPU and where from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

class A: pass

def main():
    from diffusion_pruning_b200 import HyperStructure, StructureVectorQuantizer
    from diffusion_pruning_b200 import pruning_step as PS
    from diffusion_pruning_b200.synthetic import DEPTH_ORDER
    from diffusion_pruning_b200.unet import UNet2DConditionModelGated
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    with torch.device(dev):
        unet = UNet2DConditionModelGated()
    unet.eval(); unet.freeze()
    st = unet.get_structure()
    hyper = HyperStructure(structure=st, input_dim=768, wn_flag=False, linear_bias=True).to(dev)
    quant = StructureVectorQuantizer(n_e=8, structure=st, beta=0.25, temperature=0.4, base=3, depth_order=list(DEPTH_ORDER),
                                     non_zero_width=True, resource_aware_normalization=False, optimal_transport=True).to(dev)
    quant.train(); hyper.train()
    unet.count_macs(64, 64)
    cfg = PS.PruningLossConfig()
    p_actual = PS.actual_pruning_target(unet, cfg.pruning_target)
    taps = PS.BlockTaps(unet)
    g = torch.Generator().manual_seed(100)
    Bt = 32
    batch = {"noisy_latents": torch.randn(Bt, 4, 64, 64, generator=g).to(dev), "timesteps": torch.randint(0, 1000, (Bt,), generator=g).to(dev),
             "target": torch.randn(Bt, 4, 64, 64, generator=g).to(dev), "encoder_hidden_states": torch.randn(Bt, 77, 1024, generator=g).to(dev),
             "mpnet_embeddings": torch.randn(Bt, 768, generator=g).to(dev)}
    acp = PS.alphas_cumprod().to(dev)
    def step():
        out = PS.pruning_step(unet, hyper, quant, batch, cfg, taps, p_actual, acp=acp)
        out["loss"].backward()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=50, max_shapes_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))

main()
