"""N>1 on real GPUs (NCCL over NVLink): the sharded Sinkhorn router with its marginal all-reduces
(pdm/models/vq/quantizer.py:278-300) must give the single-process assignments. Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`); skipped on a 1-GPU box."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, B_local):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import router_checks as RC
    torch.manual_seed(5)
    q = RC.make_quantizer()
    q.train()
    g = torch.Generator().manual_seed(4)
    z_all = torch.randn(B_local * world, q.vq_embed_dim, generator=g)
    u_codes = torch.rand(8, q.vq_embed_dim, generator=g)
    u_all = torch.rand(B_local * world, q.vq_embed_dim, generator=g)
    q.embedding_gs.data = q.gumbel_sigmoid_trick(q.embedding.weight.detach(), uniforms=u_codes).detach()
    draws = iter([u_all[rank * B_local:(rank + 1) * B_local]])
    q._draw_uniforms = lambda batch: next(draws)
    idx = q.get_optimal_transport_min_encoding_indices(z_all[rank * B_local:(rank + 1) * B_local].cuda())
    torch.cuda.synchronize()
    torch.save(idx.cpu(), f"{out_path}.{rank}")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_router_nccl_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    import torch.nn as nn
    world, B_local = 2, 2048
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), out, B_local), nprocs=world, join=True)
    got = torch.cat([torch.load(f"{out}.{r}") for r in range(world)])
    sys.path.insert(0, HERE)
    import router_checks as RC
    from oracle import router_oracle as R
    from oracle.structure import sd21_gate_structure
    layout = R.ArchLayout(sd21_gate_structure(), RC.DEPTH_ORDER)
    torch.manual_seed(5)
    emb = nn.Embedding(8, layout.dim)
    nn.init.orthogonal_(emb.weight)
    g = torch.Generator().manual_seed(4)
    z_all = torch.randn(B_local * world, layout.dim, generator=g)
    u_codes = torch.rand(8, layout.dim, generator=g)
    u_all = torch.rand(B_local * world, layout.dim, generator=g)
    codes = R.gumbel_sigmoid_trick(emb.weight.detach(), u_codes, layout, 0.4, 3.0)
    z_gs = R.gumbel_sigmoid_trick(z_all, u_all, layout, 0.4, 3.0)
    ref_idx, _, ref_Q = R.ot_indices(z_gs, codes, layout)
    RC._assert_assignments(got, ref_idx, ref_Q, "sharded Sinkhorn (2 ranks, NCCL)", amplification=1.0 / 0.05)
