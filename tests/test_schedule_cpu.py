"""Host-side logic of the GEMM schedules (CPU only): tile-width choice, conv boxes, tile lists, and the ragged last column
tile as the kernel computes it (`tile_mma_n` in csrc/gemm_sm100.cu, restated here)."""
import numpy as np
import pytest
import torch

from diffusion_pruning_b200 import kernels as K
from diffusion_pruning_b200 import plan as P
from diffusion_pruning_b200._lib import A_CONV3X3, A_CONV3X3_S2, A_LINEAR


def tile_mma_n(n_valid, n_store, n0, bn, geglu=False, flags=0):
    """Accumulator columns the MMAs of a tile compute (csrc/gemm_sm100.cu: tile_mma_n)."""
    if geglu:
        return bn
    w32 = (flags >> 8) & 0xFF
    width = w32 * 32 if w32 else bn
    left = (max(n_valid, n_store) - n0 + 31) // 32 * 32
    return width if left >= width else max(left, 32)


@pytest.mark.parametrize("geglu", [False, True])
def test_choose_bn_returns_a_legal_width_and_covers_every_column(geglu):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n_values = [int(rng.integers(1, 41)) * 32 for _ in range(int(rng.integers(1, 9)))]
        k = int(rng.choice([0, 320, 640, 1280, 2880, 11520]))
        bn = P.choose_bn(n_values, geglu=geglu, k=k)
        assert 32 <= bn <= 256 and bn % (64 if geglu else 32) == 0
        cols = bn // 2 if geglu else bn
        for n in n_values:  # the tiles of a bucket cover its columns, the ragged one with a legal MMA N
            tiles = (n + cols - 1) // cols
            assert tiles * cols >= n
            n_last = tile_mma_n(n, n, (tiles - 1) * bn, bn, geglu)
            assert n_last % 32 == 0 and 32 <= n_last <= bn
            if not geglu:
                assert (tiles - 1) * bn + n_last >= n


def test_choose_bn_prefers_wide_tiles_when_the_remainder_is_cheap():
    # N = 960 at K = 320: four (balanced) 256-wide tiles move less shared-memory traffic than 7.5 x 128
    assert P.choose_bn([960], k=320) == 256
    # N = 320 has no cheap split into 256 + 64: two 160-wide tiles
    assert P.choose_bn([320], k=320) == 160
    assert P.choose_bn([320], k=2880) == 160


@pytest.mark.parametrize("H,W,expect", [(64, 64, (8, 16, 1)), (32, 32, (8, 16, 1)), (16, 16, (8, 16, 1)), (96, 96, (8, 16, 1)),
                                         (48, 48, (8, 16, 1)), (8, 8, (8, 8, 2)), (24, 24, None)])
def test_stride1_convs_take_halo_boxes_where_the_image_allows(H, W, expect, monkeypatch):
    monkeypatch.delenv("APTP_CONV_HALO", raising=False)
    s = K.build_schedule([K.Segment(0, 2 * H * W, 64, 1)], 64, "cpu", mode=A_CONV3X3, Ho=H, Wo=W)
    if expect is None:
        assert s.box == K.conv_box(W, H) and s.box[0] * s.box[1] * s.box[2] == 128
    else:
        assert s.box == expect
    # every output pixel of every sample is covered exactly once by the boxes of the tile list
    bw, bh, bb = s.box
    tiles = s.tiles.numpy()
    tiles = tiles[(tiles[:, 3] & 1) == 0]
    seen = np.zeros(2 * H * W, dtype=np.int32)
    for _, m_base, n0, _ in tiles:
        if n0 != 0:
            continue
        img, rem = divmod(int(m_base), H * W)
        oy, ox = divmod(rem, W)
        for ib in range(bb):
            for y in range(bh):
                seen[(img + ib) * H * W + (oy + y) * W + ox: (img + ib) * H * W + (oy + y) * W + ox + bw] += 1
    assert (seen == 1).all()
    monkeypatch.setenv("APTP_CONV_HALO", "0")
    s0 = K.build_schedule([K.Segment(0, 2 * H * W, 64, 1)], 64, "cpu", mode=A_CONV3X3, Ho=H, Wo=W)
    assert s0.box == K.conv_box(W, H)


def test_buckets_get_balanced_column_tiles(monkeypatch):
    """N = 320 under bn = 224 -> 160 + 160 (width in flags bits 8..15), N = 192 -> one 192-wide tile; the MMAs of every tile
    cover exactly the bucket's columns."""
    monkeypatch.delenv("APTP_BALANCED_TILES", raising=False)
    segs = [K.Segment(0, 256, 320, 5), K.Segment(256, 512, 192, 5, w_row_off=320), K.Segment(512, 768, 224, 5, w_row_off=640)]
    s = K.build_schedule(segs, 224, "cpu", mode=A_LINEAR)
    t = s.tiles.numpy()
    for si, n in enumerate((320, 192, 224)):
        mine = t[t[:, 0] == si]
        n0s = sorted(set(mine[:, 2].tolist()))
        widths = [tile_mma_n(n, n, n0, 224, flags=int(mine[0, 3])) for n0 in n0s]
        assert sum(widths) >= n and all(w % 32 == 0 for w in widths)
        assert [a + w for a, w in zip(n0s[:-1], widths[:-1])] == n0s[1:]   # tiles abut
    assert sorted(set(t[t[:, 0] == 0][:, 2].tolist())) == [0, 160]
    assert sorted(set(t[t[:, 0] == 1][:, 2].tolist())) == [0]
    monkeypatch.setenv("APTP_BALANCED_TILES", "0")
    s0 = K.build_schedule(segs, 224, "cpu", mode=A_LINEAR)
    assert sorted(set(s0.tiles.numpy()[s0.tiles.numpy()[:, 0] == 0][:, 2].tolist())) == [0, 224]


def test_stride2_convs_keep_the_generic_boxes():
    s = K.build_schedule([K.Segment(0, 2 * 32 * 32, 64, 1)], 64, "cpu", mode=A_CONV3X3_S2, Ho=32, Wo=32)
    assert s.box == K.conv_box(32, 32)


def test_tile_pairs_share_bucket_and_column_block_and_odd_buckets_get_a_placeholder():
    segs = [K.Segment(0, 3 * 128, 320, 5), K.Segment(3 * 128, 3 * 128 + 200, 192, 5, w_row_off=320)]
    s = K.build_schedule(segs, 160, "cpu", mode=A_LINEAR)
    t = s.tiles.numpy().reshape(-1, 2, 4)
    assert (t[:, 0, 0] == t[:, 1, 0]).all() and (t[:, 0, 2] == t[:, 1, 2]).all()
    # bucket 0 has 3 row tiles -> one placeholder per column block; bucket 1 has 2 row tiles -> none
    ph = (t[:, :, 3] & 1).sum(1)
    assert ph[t[:, 0, 0] == 0].sum() == 2 and ph[t[:, 0, 0] == 1].sum() == 0
    assert s.n_tiles % 2 == 0


def test_uploads_of_one_state_share_device_slabs():
    """The ~500 small arrays a new assignment uploads are views into a few slabs (no per-array allocation); on CPU the
    stager simply copies."""
    a = K.upload(np.arange(12, dtype=np.int32).reshape(3, 4), "cpu")
    assert torch.equal(a, torch.arange(12, dtype=torch.int32).reshape(3, 4))
