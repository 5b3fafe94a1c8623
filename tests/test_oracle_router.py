"""CPU tests: the router / gate / hypernet / loss oracle (oracle/router_oracle.py) against golden vectors
produced by the reference's own code (tests/golden/make_goldens.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import router_oracle as R
from oracle.structure import sd21_gate_structure

G = os.path.join(os.path.dirname(__file__), "golden")
DEPTH_ORDER = [-1, -2, 0, 1, -3, -4, 2, 3, -5, -6, 4, 5, -7, 6]


@pytest.fixture(scope="module")
def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "router.npz")).items()}


@pytest.fixture(scope="module")
def layout():
    return R.ArchLayout(sd21_gate_structure(), DEPTH_ORDER)


def test_layout_totals(layout):
    # quantizer.py:44-52 on the shipped SD-2.1 structure: 70 width gates, 1606 + 14 = 1620 columns
    assert len(layout.width_list) == 70 and layout.n_width == 1606 and layout.n_depth == 14 and layout.dim == 1620
    assert layout.depth_order == [13, 12, 0, 1, 11, 10, 2, 3, 9, 8, 4, 5, 7, 6]
    assert layout.depth_indices[1] == 1606 and layout.depth_indices[-1] == 1619


@pytest.mark.parametrize("tag", ["a", "b"])
def test_train_gates_and_ot_indices(gold, layout, tag):
    z = gold["z" if tag == "a" else "z2"]
    codes_gs = R.gumbel_sigmoid_trick(gold["codebook"], gold[f"train_{tag}_u_codes"], layout, 0.4, 3.0)
    assert torch.allclose(codes_gs, gold[f"train_{tag}_codes_gs"], atol=1e-6, rtol=0)
    z_gs = R.gumbel_sigmoid_trick(z, gold[f"train_{tag}_u_z"], layout, 0.4, 3.0)
    assert torch.allclose(z_gs, gold[f"train_{tag}_z_gs"], atol=1e-6, rtol=0)
    idx, _, _ = R.ot_indices(z_gs, codes_gs, layout)
    assert torch.equal(idx, gold[f"train_{tag}_idx"])
    assert torch.allclose(codes_gs[idx], gold[f"train_{tag}_zq"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_eval_gates_and_cosine_indices(gold, layout, tag):
    z = gold["z" if tag == "a" else "z2"]
    z_gs = R.gumbel_sigmoid_trick(z, gold["eval_u"], layout, 0.4, 3.0)
    assert torch.allclose(z_gs, gold[f"eval_{tag}_z_gs"], atol=1e-6, rtol=0)
    idx, _ = R.cosine_indices(z_gs, gold["eval_embedding_gs"], layout)
    assert torch.equal(idx, gold[f"eval_{tag}_idx"])
    zq = R.hard_concrete(gold["eval_embedding_gs"][idx])
    assert torch.equal(zq, gold[f"eval_{tag}_zq"])
    assert set(np.unique(zq.numpy()).tolist()) <= {0.0, 1.0}


def test_nonzero_width_fixup_triggers(gold, layout):
    # z2 has strongly negative logits: at least one slice must have needed the +0.5 fix-up
    gs = gold["train_b_z_gs"]
    hit = 0
    for i in range(len(layout.width_list)):
        s, e = layout.width_starts[i], layout.width_starts[i + 1]
        sl = gs[:, s:e]
        only_first = (sl[:, 0] >= 0.5) & ((sl[:, 1:] >= 0.5).sum(1) == 0)
        hit += int(only_first.sum())
        assert ((sl >= 0.5).sum(1) >= 1).all(), "non_zero_width must leave >= 1 unit on"
    assert hit > 0


def test_width_depth_normalize(gold, layout):
    assert torch.allclose(R.width_depth_normalize(gold["eval_a_z_gs"], layout), gold["norm_z_gs_a"], atol=1e-7)
    assert torch.allclose(R.width_depth_normalize(gold["soft"], layout), gold["norm_soft"], atol=1e-7)


def test_uniform_draw_order_matches_flat_draw(layout):
    # the per-slice draws are equivalent to one flat draw sliced in the same order
    torch.manual_seed(11)
    u = R.draw_uniforms(layout, 4, fixed_seed=False)
    torch.manual_seed(11)
    ud = torch.rand(4, layout.n_depth)
    assert torch.equal(u[:, layout.n_width:], ud)
    uw0 = torch.rand(4, layout.width_list[0])
    assert torch.equal(u[:, :layout.width_list[0]], uw0)


def test_hypernet_and_split():
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "hypernet.npz")).items()}
    layout = R.ArchLayout(sd21_gate_structure(), DEPTH_ORDER)
    sizes = layout.width_list + [layout.n_depth]
    ws = list(torch.split(g["weight"], sizes, 0))
    bs = list(torch.split(g["bias"], sizes, 0))
    y = R.hypernet_forward(g["x"], ws, bs)
    assert torch.allclose(y, g["y"], atol=1e-5)
    assert int(g["n_width"]) == 70 and int(g["n_depth"]) == 14
    s3 = layout.width_starts[3]
    assert torch.equal(g["y"][:, s3:s3 + layout.width_list[3]], g["width_3"])
    assert torch.equal(g["y"][:, layout.n_width + 5], g["depth_5"])


def test_losses_and_gates():
    from oracle import unet_oracle as U
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "losses_gates.npz")).items()}
    assert torch.allclose(R.contrastive_loss(g["prompt"], g["arch"], 0.03, 0.03), g["contrastive"], atol=1e-6)
    assert torch.allclose(R.resource_loss(torch.tensor(0.8), 0.6), g["resource_hi"])
    assert torch.allclose(R.resource_loss(torch.tensor(0.4), 0.6), g["resource_lo"])
    gw = torch.tensor([[1., 0., 0.5, 1.], [0., 1., 1., 0.25]])
    assert torch.equal(U.width_gate(g["wg_x"], gw), g["wg_y"])  # includes the CFG batch-doubling repeat
    assert torch.equal(U.depth_gate(g["wg_x"], g["dg_y_in"], torch.tensor([0.25, 1.0])), g["dg_out"])
    assert torch.equal(U.linear_width_gate(g["lg_x"], gw), g["lg_y"])
