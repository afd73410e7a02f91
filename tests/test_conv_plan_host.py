"""Host logic of conv_tc_build (no GPU): the schedule the library chooses per layer — tile width, CTA pairs, persistent grid,
small-batch fill, row-patch mode — through mf_op_conv_tc_plan, a pure function of the shape and the knobs."""
import ctypes

import pytest

from medfusion_b200 import _lib


def plan(N, H, W, C0, C1, Cout, k, stride=1, up2=0, sms=148):
    lib = _lib.load()
    out = (ctypes.c_int * 8)()
    _lib.check(lib.mf_op_conv_tc_plan(N, H, W, C0, C1, Cout, k, stride, up2, sms, out), "conv_tc_plan")
    keys = ("supported", "block_n", "cta_group", "groups", "row3", "num_tiles", "nkb", "m_groups")
    return dict(zip(keys, list(out)))


def test_headline_batch_keeps_one_pair_per_tile():
    """B = 64 (BASELINE configs[1]): 256-wide tiles on CTA pairs, one persistent pair per 2 SMs, no K sharing where the tiles
    fill at least half the pairs (the r1 experiment that shared K at B = 64 was slower under the power cap)."""
    p = plan(64, 32, 32, 256, 0, 256, 3)
    assert p == dict(supported=1, block_n=256, cta_group=2, groups=74, row3=0, num_tiles=256, nkb=36, m_groups=256)
    p = plan(64, 8, 8, 1024, 1024, 1024, 3)                    # concat conv of the 8x8 level: 64 tiles of 256 x 256
    assert (p["block_n"], p["num_tiles"], p["groups"], p["nkb"]) == (256, 64, 64, 288)
    p = plan(64, 16, 16, 512, 0, 512, 3)
    assert (p["block_n"], p["num_tiles"], p["groups"]) == (256, 128, 74)


@pytest.mark.parametrize("B", [1, 2, 4, 8, 16])
def test_small_batch_fill_uses_most_of_the_gpu(B):
    """scripts/sample.py batches: every 3x3 layer of the canonical UNet gets a grid of at least half the SM pairs (it was
    min(tiles, pairs): 4 pairs of 74 at B = 4 on the 8x8 level), each group keeps >= 8 K blocks, tiles stay >= 64 wide."""
    lib = _lib.load()
    try:
        for (H, C0, C1, Cout) in ((32, 256, 0, 256), (32, 256, 256, 256), (16, 512, 0, 512), (8, 1024, 0, 1024),
                                  (8, 1024, 1024, 1024)):
            lib.mf_set_split_fill(8)
            p = plan(B, H, H, C0, C1, Cout, 3)
            assert p["supported"] == 1 and p["block_n"] in (64, 128, 256)
            units = p["num_tiles"] * p["nkb"]
            assert p["groups"] >= min(37, units // 8), (B, H, p)
            assert p["groups"] <= 74 * (2 // p["cta_group"]) and p["groups"] * 8 <= max(units, 8 * p["num_tiles"]) + 8 * 74
            lib.mf_set_split_fill(0)
            q = plan(B, H, H, C0, C1, Cout, 3)
            assert q["block_n"] == 256 and q["groups"] == min(q["num_tiles"], 74 * (2 // q["cta_group"]))   # the old plan
            assert p["groups"] >= q["groups"]
    finally:
        lib.mf_set_split_fill(8)


def test_row_patch_mode_only_for_row_tiled_narrow_layers():
    """Row-patch staging applies to 3x3 stride-1 layers whose tile is 128 pixels of one image row and at most 128 channels
    wide (the VAE's 128x128 / 256x256 levels, and the phases of the folded up-conv that ends there), nowhere else."""
    lib = _lib.load()
    assert plan(64, 256, 256, 64, 0, 64, 3)["row3"] == 1
    assert plan(64, 128, 128, 128, 0, 128, 3)["row3"] == 1
    assert plan(64, 128, 128, 128, 0, 64, 3, up2=1)["row3"] == 1          # 128x128 -> 256x256 folded up-conv
    assert plan(64, 64, 64, 256, 0, 128, 3, up2=1)["row3"] == 0           # tiles of two image rows
    assert plan(64, 128, 128, 256, 0, 256, 3)["row3"] == 0                # 256-wide tiles fit under the L2 -> SM fabric
    assert plan(64, 32, 32, 256, 0, 256, 3)["row3"] == 0                  # UNet levels: tiles span several rows
    assert plan(64, 256, 256, 64, 0, 64, 1)["row3"] == 0                  # 1x1
    assert plan(64, 256, 256, 64, 0, 64, 3, stride=2)["row3"] == 0
    try:
        lib.mf_set_row_patch(0)
        assert plan(64, 256, 256, 64, 0, 64, 3)["row3"] == 0
    finally:
        lib.mf_set_row_patch(1)


def test_unsupported_shapes_are_reported_not_planned():
    assert plan(2, 32, 32, 96, 0, 256, 3)["supported"] == 0               # channels not a multiple of 64 -> SIMT path
    assert plan(2, 24, 24, 256, 0, 256, 3)["supported"] == 0              # 24 x 24 does not tile into 128-pixel boxes
    assert plan(2, 33, 33, 256, 0, 256, 3, stride=2)["supported"] == 0
