"""Where do the allocations of a re-structured (fresh assignment) forward go?  python tools/prof_restructure2.py"""
import sys, time, collections
sys.path.insert(0, '.')
import torch
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
orig = torch.empty
acc = collections.defaultdict(lambda: [0, 0.0])
def timed_empty(*a, **k):
    t0 = time.perf_counter()
    r = orig(*a, **k)
    dt = time.perf_counter() - t0
    key = 'pinned' if k.get('pin_memory') else ('cpu' if r.device.type == 'cpu' else ('>=64MB' if r.numel() * r.element_size() >= 64 << 20 else ('>=1MB' if r.numel() * r.element_size() >= 1 << 20 else '<1MB')))
    acc[key][0] += 1; acc[key][1] += dt
    return r
import gc
if len(sys.argv) > 1: gc.disable()
for rep in range(5):
    arch = fresh()
    torch.cuda.synchronize()
    acc.clear()
    torch.empty = timed_empty
    t0 = time.perf_counter()
    model.set_structure(split_arch(arch, st))
    with torch.no_grad(): model(s, tt, c)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    torch.empty = orig
    print(f'rep{rep} restructure {dt:.1f} ms;', {k: (v[0], round(v[1] * 1e3, 2)) for k, v in acc.items()})
    print('   reserved GB', torch.cuda.memory_reserved() / 2**30, 'allocated GB', torch.cuda.memory_allocated() / 2**30)
