"""GPU parity: the CUDA U-Net path vs the fp32 CPU oracle on identical seeded weights and inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _assert(res):
    import unet_checks as U
    max_abs, cos = res
    assert max_abs <= U.MAX_ABS_TOL and cos >= U.COS_TOL, (max_abs, cos)


def test_forward_is_bit_reproducible():
    import unet_checks as U
    assert U.check_deterministic()


def test_tiny_all_ones_equals_ungated():
    import unet_checks as U
    _assert(U.check_all_ones())


@pytest.mark.parametrize("beta_std", [0.0, 0.1])
def test_tiny_hard_mixed_experts(beta_std):
    import unet_checks as U
    _assert(U.check_hard(beta_std=beta_std))


def test_tiny_soft_gates():
    import unet_checks as U
    _assert(U.check_soft())


def test_tiny_cfg_batch_doubling():
    import unet_checks as U
    _assert(U.check_cfg_doubling())


def test_full_sd21_hard_b2_h32():
    import unet_checks as U
    _assert(U.check_hard(tiny=False, B=2, H=32, code_ids=(0, 3)))


def test_tiny_pruned_static_expert_matches_pruned_oracle():
    import unet_checks as U
    res, gap = U.check_pruned_expert()
    _assert(res)
    assert gap > U.MAX_ABS_TOL, gap  # the case really separates prune() from gate semantics


def test_full_sd21_hard_b8_h64_eight_codes():
    """SURVEY 8(d) config 2 parity slice: full-size SD-2.1 weights, 64x64 latents, 8 samples on 8 distinct codes (width +
    depth gating), GroupNorm beta != 0; bf16 CUDA path vs the fp32 CPU oracle."""
    import unet_checks as U
    _assert(U.check_hard(tiny=False, B=8, H=64, code_ids=(0, 1, 2, 3, 4, 5, 6, 7), beta_std=0.1))


@pytest.mark.parametrize("B,H,code_ids", [
    (1, 32, (5,)),                 # a single sample / single expert
    (5, 16, (2, 2, 2, 2, 2)),      # odd batch, every sample on the same expert (one bucket)
    (2, 96, (1, 6)),               # sampling resolution: 9216 / 2304 / 576 / 144 tokens (ragged last key tiles)
    (3, 24, (0, 7, 4)),            # 24x24: 8x8x2 conv boxes with a partial batch box, 576 / 144 / 36 / 9 tokens
])
def test_tiny_hard_edge_shapes(B, H, code_ids):
    import unet_checks as U
    _assert(U.check_hard(B=B, H=H, code_ids=code_ids, beta_std=0.1))


@pytest.mark.parametrize("case", ["hard", "soft", "cfg", "ones"])
def test_tiny_vs_reference_executed_goldens(case):
    """CUDA path vs outputs recorded from the reference's own forward (tests/golden/unet_ref.npz)."""
    import unet_checks as U
    res, tap_cos = U.check_vs_reference_golden(case)
    _assert(res)
    assert all(c >= 0.9995 for c in tap_cos), tap_cos  # the nine hooked block outputs (bf16 taps)


def test_engine_caches_stay_bounded_over_many_assignments():
    """schedules / packs must not grow without bound when every batch brings a new prompt -> expert assignment
    (pruning_pipelines.py:757-759 calls set_structure per batch), and CUDA graphs must survive the eviction."""
    import unet_checks as U
    from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
    model, oracle = U.build_pair(True, beta_std=0.1)
    st = model.get_structure()
    codes = synthetic_codes(st, 8)
    sample, t, ctx = U.inputs(4, 16, model.config["cross_attention_dim"])
    g = torch.Generator().manual_seed(0)
    first = None
    assign0 = [0, 3, 3, 7]
    sizes = []
    for it in range(40):
        assign = assign0 if it % 13 < 3 else torch.randint(0, 8, (4,), generator=g).tolist()
        model.set_structure(split_arch(codes[assign].clone().cuda(), st))
        with torch.no_grad():
            y = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
        if assign == assign0:
            first = y.clone() if first is None else first
            assert torch.equal(y, first), "a replayed / re-built state must reproduce the same bits"
        eng = model._engine
        sizes.append((len(eng.sched), len(eng.expert), len(eng.graphs), len(eng.graph_seen)))
    assert len(eng._states) <= eng.MAX_STATES and len(eng._esets) <= eng.MAX_ESETS
    assert max(s[0] for s in sizes[20:]) <= max(s[0] for s in sizes[:20]) * 1.5 + 50, sizes[-1]
    assert sizes[-1][2] <= 4


def test_new_assignment_with_same_bucket_sizes_reuses_the_cached_graph():
    """Cached schedules / CUDA graph are keyed on bucket SIZES: a different prompt -> expert assignment with the same sizes
    must only rewrite the permutation buffers and still match the oracle (and the graph must really be reused)."""
    import unet_checks as U
    from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes
    model, oracle = U.build_pair(True, beta_std=0.1)
    st = model.get_structure()
    codes = synthetic_codes(st, 8)
    sample, t, ctx = U.inputs(4, 16, model.config["cross_attention_dim"])
    for it, assign in enumerate(([0, 3, 3, 7], [0, 3, 3, 7], [0, 3, 3, 7], [3, 0, 7, 3], [7, 3, 0, 3], [3, 3, 7, 0])):
        arch = codes[assign]
        model.set_structure(split_arch(arch.clone().cuda(), st))
        with torch.no_grad():
            got = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
        oracle.set_structure(split_arch(arch.clone(), st))
        with torch.no_grad():
            ref = oracle(sample, t, ctx)
        _assert(U.metrics(got, ref))
    eng = model._engine
    assert len(eng.graphs) == 1 and len(eng._states) == 1, (len(eng.graphs), len(eng._states))
