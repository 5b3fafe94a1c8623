"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):  python tests/golden/make_goldens.py
The reference sources are loaded in place through oracle/ref_shim (a stub `diffusers`); nothing is
copied. The resulting .npz files are small, committed, and are what the CPU tests pin the oracle
(oracle/router_oracle.py, oracle/unet_oracle.py) against; the GPU tests then compare the CUDA path with
the oracle and with these files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle.ref_shim import load_router_reference  # noqa: E402
from oracle.router_oracle import ArchLayout, draw_uniforms  # noqa: E402
from oracle.unet_oracle import GatedUNetOracle  # noqa: E402

DEPTH_ORDER = [-1, -2, 0, 1, -3, -4, 2, 3, -5, -6, 4, 5, -7, 6]  # configs/pruning/sd-2-1_cc3m.yaml:38


def sd21_structure():
    return GatedUNetOracle.__new__(GatedUNetOracle)  # placeholder (unused)


def router_goldens():
    ref = load_router_reference()
    from oracle.unet_oracle import UNetConfig
    torch.manual_seed(0)
    # the real SD-2.1 gate structure (70 width gates / 1606 / 14 depth), without allocating weights
    from oracle.structure import sd21_gate_structure
    structure = sd21_gate_structure()
    layout = ArchLayout(structure, DEPTH_ORDER)
    assert layout.dim == 1620 and len(layout.width_list) == 70 and layout.n_depth == 14

    B, K = 96, 8
    torch.manual_seed(5)
    vq = ref["quantizer"].StructureVectorQuantizer(
        n_e=K, structure=structure, beta=0.25, temperature=0.4, base=3, depth_order=list(DEPTH_ORDER),
        non_zero_width=True, resource_aware_normalization=False, optimal_transport=True)
    assert vq.vq_embed_dim == 1620
    g = torch.Generator().manual_seed(4)
    z = torch.randn(B, 1620, generator=g)
    # a second input with strongly negative logits so the non_zero_width fix-up and pruning trigger
    z2 = torch.randn(B, 1620, generator=g) * 3.0 - 4.0
    codebook = vq.embedding.weight.detach().clone()

    out = {"z": z.numpy(), "z2": z2.numpy(), "codebook": codebook.numpy()}
    # ---- train mode: gumbel codes -> Sinkhorn OT indices (quantizer.py:140-151) ----
    vq.train()
    for tag, zz, seed in (("a", z, 123), ("b", z2, 321)):
        torch.manual_seed(seed)
        z_q, (_, _, idx) = vq(zz)
        torch.manual_seed(seed)
        u_codes = draw_uniforms(layout, K, fixed_seed=False)
        u_z = draw_uniforms(layout, B, fixed_seed=False)
        out[f"train_{tag}_u_codes"] = u_codes.numpy()
        out[f"train_{tag}_u_z"] = u_z.numpy()
        out[f"train_{tag}_codes_gs"] = vq.embedding_gs.detach().numpy().copy()
        out[f"train_{tag}_idx"] = idx.numpy()
        out[f"train_{tag}_zq"] = z_q.detach().numpy()
        # the gated input itself (what gumbel_sigmoid_trick returned for z) -- replay with the same draws
        torch.manual_seed(seed)
        _ = draw_uniforms(layout, K, fixed_seed=False)
        out[f"train_{tag}_z_gs"] = vq.gumbel_sigmoid_trick(zz).detach().numpy()
    # ---- eval mode: stored embedding_gs, cosine argmax, hard_concrete (quantizer.py:147-167) ----
    vq.eval()
    for tag, zz in (("a", z), ("b", z2)):
        z_q, (_, _, idx) = vq(zz)
        out[f"eval_{tag}_idx"] = idx.numpy()
        out[f"eval_{tag}_zq"] = z_q.detach().numpy()
        out[f"eval_{tag}_z_gs"] = vq.gumbel_sigmoid_trick(zz).detach().numpy()
    out["eval_embedding_gs"] = vq.embedding_gs.detach().numpy().copy()
    out["eval_u"] = draw_uniforms(layout, B, fixed_seed=True).numpy()
    out["norm_z_gs_a"] = vq.width_depth_normalize(torch.from_numpy(out["eval_a_z_gs"])).numpy()
    # random-ish soft vectors through width_depth_normalize (exercises the product-rule columns)
    soft = torch.rand(B, 1620, generator=g)
    out["soft"] = soft.numpy()
    out["norm_soft"] = vq.width_depth_normalize(soft).numpy()
    np.savez_compressed(os.path.join(OUT, "router.npz"), **out)

    # ---- hypernet (hypernet.py:28-101) with a small input dim to keep the fixture small ----
    torch.manual_seed(7)
    hn = ref["hypernet"].HyperStructure(structure=structure, input_dim=32, wn_flag=False, linear_bias=True)
    x = torch.randn(16, 32, generator=g)
    y = hn(x)
    hout = {"x": x.numpy(), "y": y.detach().numpy(),
            "weight": torch.cat([l.weight.detach() for l in hn.mh_fc], 0).numpy(),
            "bias": torch.cat([l.bias.detach() for l in hn.mh_fc], 0).numpy()}
    sv = hn.transform_structure_vector(y.detach())
    hout["n_width"] = np.array(len(sv["width"]))
    hout["n_depth"] = np.array(len(sv["depth"]))
    hout["width_3"] = sv["width"][3].numpy()
    hout["depth_5"] = sv["depth"][5].numpy()
    np.savez_compressed(os.path.join(OUT, "hypernet.npz"), **hout)

    # ---- losses ----
    cl = ref["contrastive_loss"].ContrastiveLoss(arch_vector_temperature=0.03, prompt_embedding_temperature=0.03)
    pe = torch.randn(24, 48, generator=g)
    av = torch.rand(24, 1620, generator=g)
    lout = {"prompt": pe.numpy(), "arch": av.numpy(), "contrastive": cl(pe, av).numpy()}
    rl = ref["resource_loss"].ResourceLoss(p=0.6, loss_type="log")
    lout["resource_hi"] = rl(torch.tensor(0.8)).numpy()
    lout["resource_lo"] = rl(torch.tensor(0.4)).numpy()
    # gates.py forward semantics incl. CFG batch doubling
    gates = ref["gates"]
    wg = gates.WidthGate(4)
    wg.set_structure_value(torch.tensor([[1., 0., 0.5, 1.], [0., 1., 1., 0.25]]))
    xg = torch.randn(4, 8, 3, 3, generator=g)
    lout["wg_x"] = xg.numpy()
    lout["wg_y"] = wg(xg).numpy()
    dg = gates.DepthGate(1)
    dg.set_structure_value(torch.tensor([0.25, 1.0]))
    yg = torch.randn(4, 8, 3, 3, generator=g)
    lout["dg_y_in"] = yg.numpy()
    lout["dg_out"] = dg((xg, yg)).numpy()
    lg = gates.LinearWidthGate(4)
    lg.set_structure_value(torch.tensor([[1., 0., 0.5, 1.], [0., 1., 1., 0.25]]))
    xl = torch.randn(4, 5, 8, generator=g)
    lout["lg_x"] = xl.numpy()
    lout["lg_y"] = lg(xl).numpy()
    np.savez_compressed(os.path.join(OUT, "losses_gates.npz"), **lout)
    print("router goldens written")


if __name__ == "__main__":
    which = sys.argv[1:] or ["router", "unet"]
    if "router" in which:
        router_goldens()
    if "unet" in which:
        sys.path.insert(0, OUT)
        from make_unet_goldens import unet_goldens
        unet_goldens()
