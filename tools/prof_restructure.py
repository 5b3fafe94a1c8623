import sys, time, torch, cProfile, pstats
sys.path.insert(0, '/root/repo')
import bench
from diffusion_pruning_b200.synthetic import split_arch
dev = torch.device('cuda')
model, codes, assign, sample, ctx, t = bench.make_workload(dev, 1234)
s, c, tt = sample.to(dev), ctx.to(dev), t.to(dev)
with torch.no_grad():
    for _ in range(3): model(s, tt, c)
torch.cuda.synchronize()
st = model.get_structure()
g = torch.Generator().manual_seed(5)
def fresh():
    a = torch.randint(0, 8, (64,), generator=g)
    return codes[a].to(dev)
arch = fresh()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
model.set_structure(split_arch(arch, st))
with torch.no_grad(): model(s, tt, c)
torch.cuda.synchronize()
print('restructure ms', (time.perf_counter() - t0) * 1e3)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
