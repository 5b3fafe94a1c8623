"""world_size-2 `gloo` tests (CPU) of the N>1 HOST logic: the sharded Sinkhorn phase protocol with its four
marginal all-reduces (pdm/models/vq/quantizer.py:278-300) and the all-gather + local-slice re-insertion of the
contrastive loss (pdm/training/trainer.py:1153-1160). The CUDA kernels are replaced by the torch emulation of
their contract (tests/cpu_kernel_sim.py); the real kernels run under NCCL in tests/test_multigpu.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEPTH_ORDER = [-1, -2, 0, 1, -3, -4, 2, 3, -5, -6, 4, 5, -7, 6]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, B_local):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_kernel_sim
    cpu_kernel_sim.install_router(setattr)
    from diffusion_pruning_b200 import StructureVectorQuantizer
    from diffusion_pruning_b200 import pruning_step as PS
    from oracle.structure import sd21_gate_structure
    torch.manual_seed(5)
    q = StructureVectorQuantizer(n_e=8, structure=sd21_gate_structure(), beta=0.25, temperature=0.4, base=3,
                                 depth_order=list(DEPTH_ORDER), non_zero_width=True,
                                 resource_aware_normalization=False, optimal_transport=True)
    q.train()
    g = torch.Generator().manual_seed(4)
    z_all = torch.randn(B_local * world, q.vq_embed_dim, generator=g)
    u_codes = torch.rand(8, q.vq_embed_dim, generator=g)
    u_all = torch.rand(B_local * world, q.vq_embed_dim, generator=g)
    z = z_all[rank * B_local:(rank + 1) * B_local]
    u = u_all[rank * B_local:(rank + 1) * B_local]
    # the same codes on every rank (DDP keeps the codebook in sync), this rank's shard of prompts
    q.embedding_gs.data = q.gumbel_sigmoid_trick(q.embedding.weight.detach(), uniforms=u_codes).detach()
    draws = iter([u])
    q._draw_uniforms = lambda batch: next(draws)
    idx = q.get_optimal_transport_min_encoding_indices(z)
    # contrastive all-gather path of the train step
    text = torch.randn(B_local * world, 16, generator=g)[rank * B_local:(rank + 1) * B_local].requires_grad_(True)
    arch = torch.rand(B_local * world, 40, generator=g)[rank * B_local:(rank + 1) * B_local].requires_grad_(True)
    tl = [torch.zeros_like(text) for _ in range(world)]
    al = [torch.zeros_like(arch) for _ in range(world)]
    dist.all_gather(tl, text.detach())
    dist.all_gather(al, arch.detach())
    tl[rank], al[rank] = text, arch
    loss, _ = PS.contrastive_loss(torch.cat(tl), torch.cat(al), 0.03, 0.03)
    loss.backward()
    torch.save({"idx": idx, "loss": loss.detach(), "darch": arch.grad}, f"{out_path}.{rank}")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_sinkhorn_and_contrastive_gather_match_single_process(tmp_path):
    world, B_local = 2, 24
    port = _free_port()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, port, out, B_local), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    # single-process oracle on the concatenated batch
    sys.path.insert(0, ROOT)
    from oracle import router_oracle as R
    from oracle.structure import sd21_gate_structure
    layout = R.ArchLayout(sd21_gate_structure(), DEPTH_ORDER)
    torch.manual_seed(5)
    import torch.nn as nn
    emb = nn.Embedding(8, layout.dim)
    nn.init.orthogonal_(emb.weight)
    g = torch.Generator().manual_seed(4)
    z_all = torch.randn(B_local * world, layout.dim, generator=g)
    u_codes = torch.rand(8, layout.dim, generator=g)
    u_all = torch.rand(B_local * world, layout.dim, generator=g)
    codes = R.gumbel_sigmoid_trick(emb.weight.detach(), u_codes, layout, 0.4, 3.0)
    z_gs = R.gumbel_sigmoid_trick(z_all, u_all, layout, 0.4, 3.0)
    ref_idx, _, _ = R.ot_indices(z_gs, codes, layout)
    got_idx = torch.cat([r["idx"] for r in res])
    assert torch.equal(got_idx, ref_idx), "sharded Sinkhorn assignments differ from the single-process reference"
    text = torch.randn(B_local * world, 16, generator=g)
    arch = torch.rand(B_local * world, 40, generator=g).requires_grad_(True)
    ref_loss = R.contrastive_loss(text, arch, 0.03, 0.03)
    ref_loss.backward()
    for r in range(world):
        assert torch.allclose(res[r]["loss"], ref_loss.detach(), rtol=1e-5, atol=1e-7)
        assert torch.allclose(res[r]["darch"], arch.grad[r * B_local:(r + 1) * B_local], rtol=1e-4, atol=1e-7)


def _sampling_worker(rank, world, port, out_path, P):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_pruning_b200 import sampling as S
    g = torch.Generator().manual_seed(50 + rank)
    prompt = torch.randn(P, 8, generator=g)
    latents = torch.randn(P, 4, 6, 6, generator=g)
    cond = torch.randn(P, 5, 16, generator=g)
    uncond = torch.randn(P, 5, 16, generator=g)
    idx = torch.randint(0, 8, (P,), generator=g)          # code index of every local prompt
    arch = torch.nn.functional.one_hot(idx, 8).float()    # stands in for the [P, 1620] hard architecture vectors

    def fake_route(hyper_net, quantizer, prompt_embeddings):
        return arch, idx

    def fake_denoise(unet, hyper_net, arch_vectors, lat, c, u, num_inference_steps=25, guidance_scale=7.5, acp=None,
                     prediction_type="v_prediction"):
        # every prompt that arrives here must belong to an expert this rank owns, with ITS OWN conditioning rows
        codes = arch_vectors.argmax(dim=1)
        assert bool((codes % world == rank).all()), (rank, codes)
        tag = c.mean(dim=(1, 2)) + 2.0 * u.mean(dim=(1, 2))
        return lat * 2.0 + tag[:, None, None, None] + 1000.0 * rank + 10.0 * codes[:, None, None, None].float()

    S.route_prompts, S.denoise = fake_route, fake_denoise
    import types
    out, got_idx = S.routed_sampling(None, None, types.SimpleNamespace(n_e=8), prompt, latents, cond, uncond,
                                     dispatch="expert_mod")
    tag = cond.mean(dim=(1, 2)) + 2.0 * uncond.mean(dim=(1, 2))
    want = latents * 2.0 + tag[:, None, None, None] + 1000.0 * (idx % world)[:, None, None, None].float() + \
        10.0 * idx[:, None, None, None].float()
    torch.save({"ok": bool(torch.allclose(out, want, rtol=0, atol=1e-5)), "idx_ok": bool(torch.equal(got_idx, idx)),
                "owners": torch.bincount(idx % world, minlength=world)}, f"{out_path}.{rank}")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_routed_sampling_dispatch_and_return_over_all_to_all(tmp_path, world):
    """BASELINE configs[3] host logic (sampling.routed_sampling): prompts go to the rank that owns their expert (code %
    world), are processed there with their own latents / text rows, and come back to the asking rank in the caller's
    order -- uneven expert load, variable split sizes. `denoise` is replaced by a tagging function."""
    out = str(tmp_path / "res")
    mp.spawn(_sampling_worker, args=(world, _free_port(), out, 11), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    assert all(r["ok"] for r in res) and all(r["idx_ok"] for r in res), res


def _balanced_worker(rank, world, port, out_path, P):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import types
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_pruning_b200 import sampling as S
    g = torch.Generator().manual_seed(70 + rank)
    latents = torch.randn(P, 4, 6, 6, generator=g)
    cond = torch.randn(P, 5, 16, generator=g)
    uncond = torch.randn(P, 5, 16, generator=g)
    # skewed routing (eval cosine argmax is unbalanced): ~70 % of the prompts hit expert 2
    idx = torch.where(torch.rand(P, generator=g) < 0.7, torch.full((P,), 2), torch.randint(0, 8, (P,), generator=g))
    arch = torch.nn.functional.one_hot(idx, 8).float()
    load = {}

    def fake_route(hyper_net, quantizer, prompt_embeddings):
        return arch, idx

    def fake_denoise(unet, hyper_net, arch_vectors, lat, c, u, num_inference_steps=25, guidance_scale=7.5, acp=None,
                     prediction_type="v_prediction"):
        codes = arch_vectors.argmax(dim=1)
        load["n"], load["experts"] = lat.shape[0], sorted(set(codes.tolist()))
        tag = c.mean(dim=(1, 2)) + 2.0 * u.mean(dim=(1, 2))
        return lat * 2.0 + tag[:, None, None, None] + 10.0 * codes[:, None, None, None].float()

    S.route_prompts, S.denoise = fake_route, fake_denoise
    out, got_idx = S.routed_sampling(None, None, types.SimpleNamespace(n_e=8), torch.zeros(P, 8), latents, cond, uncond)
    tag = cond.mean(dim=(1, 2)) + 2.0 * uncond.mean(dim=(1, 2))
    want = latents * 2.0 + tag[:, None, None, None] + 10.0 * idx[:, None, None, None].float()
    mod_load = torch.bincount(idx % world, minlength=world)
    torch.save({"ok": bool(torch.allclose(out, want, rtol=0, atol=1e-5)), "load": load.get("n", 0),
                "experts": load.get("experts", []), "mod_load": mod_load}, f"{out_path}.{rank}")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_balanced_dispatch_under_skewed_routing(tmp_path, world):
    """sampling.plan_dispatch("balanced"): with ~70 % of the prompts on one expert no rank serves more than
    ceil(total / world) prompts (the hot expert is replicated over several ranks), results still come back to the asking
    rank in the caller's order; the round-1 rule (expert % world) would have piled them on one rank."""
    P = 23
    out = str(tmp_path / "res")
    mp.spawn(_balanced_worker, args=(world, _free_port(), out, P), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    assert all(r["ok"] for r in res), res
    total = P * world
    cap = (total + world - 1) // world
    assert sum(r["load"] for r in res) == total and max(r["load"] for r in res) <= cap, [r["load"] for r in res]
    mod = sum(r["mod_load"] for r in res)
    assert int(mod.max()) > cap, "the case must really be skewed for expert % world"
