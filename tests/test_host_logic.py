"""Host-side planning logic that needs no GPU: tile / geometry policies of fullbatchtraining_b200.ops and the group
scheduling of the engine.  The policies must depend on ONE microbatch's problem only (never on the number of groups in
a launch): that is what makes results independent of the group count."""
import inspect

import pytest

from fullbatchtraining_b200 import engine, ops


def test_pixel_tile_covers_128_pixels():
    for h, w in [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (64, 64)]:
        tw, th, tn = ops.pixel_tile(h, w)
        assert tw == w and tw * th * tn == 128 and (tn == 1 or th == h)
    with pytest.raises(RuntimeError):
        ops.pixel_tile(24, 24)


def test_tiles_per_group_requires_whole_boxes():
    assert ops.tiles_per_group(128, 32, ops.pixel_tile(32, 32)) == 1024
    assert ops.tiles_per_group(128, 16, ops.pixel_tile(16, 16)) == 256
    assert ops.tiles_per_group(128, 8, ops.pixel_tile(8, 8)) == 64
    assert ops.tiles_per_group(128, 4, ops.pixel_tile(4, 4)) == 16
    with pytest.raises(RuntimeError):
        ops.tiles_per_group(12, 4, ops.pixel_tile(4, 4))  # 8 images per box on 4x4 maps


def test_choose_n_tile_follows_the_cost_model():
    # 64-channel stage: only 64 divides
    assert ops.choose_n_tile(1024, 64, 2, 2) == 64
    # 4x4 stage, few M tiles: 64-wide stacked tiles fill more SMs than 128-wide ones
    assert ops.choose_n_tile(16, 512, 2, 2) == 64
    # 8x8 stage: 64 M tiles x 2 N tiles of 128 = one wave
    assert ops.choose_n_tile(64, 256, 2, 2) == 128
    for m, n in [(1, 64), (7, 192), (300, 2048)]:
        assert n % ops.choose_n_tile(m, n) == 0


def test_wgrad_policies():
    assert ops.Conv2dPlan._wgrad_halo(3, 1, (32, 4, 1), 2)
    assert ops.Conv2dPlan._wgrad_halo(3, 1, (16, 8, 1), 2)
    assert not ops.Conv2dPlan._wgrad_halo(3, 1, (8, 8, 2), 2)   # tiles span two images
    assert not ops.Conv2dPlan._wgrad_halo(1, 1, (32, 4, 1), 2)
    assert not ops.Conv2dPlan._wgrad_halo(3, 2, (16, 8, 1), 2)
    assert ops.Conv2dPlan._slots_per_cta(9, 2) == 3 and ops.Conv2dPlan._slots_per_cta(72, 2) == 4
    assert ops.Conv2dPlan._slots_per_cta(2, 2) == 2 and ops.Conv2dPlan._slots_per_cta(16, 1) == 8
    # split-K: POLICY_GROUPS groups fill one wave; never more splits than pixel blocks; no split on the wide stages
    s = ops.Conv2dPlan.wgrad_splits(1024, 1, 3)
    assert 1 <= s <= 1024 and 3 * s * ops.POLICY_GROUPS <= ops.NUM_SMS
    assert ops.Conv2dPlan.wgrad_splits(16, 4, 18) == 1
    assert ops.Conv2dPlan.wgrad_splits(2, 1, 1) == 2
    # the policy has no access to the number of groups of a launch
    assert "ng" not in inspect.signature(ops.Conv2dPlan.wgrad_splits).parameters
    assert "ng" not in inspect.signature(ops.choose_n_tile).parameters


def test_launch_sizes_are_balanced():
    # as few launches as G allows, sizes equal up to one, loader order preserved by the caller
    assert engine.launch_sizes(390, 8) == [8] * 47 + [7] * 2
    assert engine.launch_sizes(49, 8) == [7] * 7
    assert engine.launch_sizes(48, 8) == [8] * 6
    assert engine.launch_sizes(15, 8) == [8, 7]
    assert engine.launch_sizes(3, 8) == [3]
    assert engine.launch_sizes(0, 8) == []
    for count in range(1, 200):
        for G in (1, 2, 5, 8, 16):
            sizes = engine.launch_sizes(count, G)
            assert sum(sizes) == count and max(sizes) <= G and max(sizes) - min(sizes) <= 1
            assert len(sizes) == -(-count // G)


def test_default_groups(monkeypatch):
    monkeypatch.delenv("FB_GROUPS", raising=False)
    assert engine.default_groups(128) == 8
    assert engine.default_groups(32) == 8
    assert engine.default_groups(2048) == 1
    monkeypatch.setenv("FB_GROUPS", "3")
    assert engine.default_groups(128) == 3
    monkeypatch.setenv("FB_GROUPS", "99")
    assert engine.default_groups(128) == 16


def test_scalar_slots_do_not_overlap():
    singles = [engine.S_LOSS, engine.S_CORRECT, engine.S_CF, engine.S_CLIPPED, engine.S_GNORM, engine.S_PNORM]
    assert len(set(singles)) == len(singles) and max(singles) < 16
    bases = [engine.S_N2G, engine.S_EPSG, engine.S_LOSSG, engine.S_CORRG, engine.S_LOSS2G, engine.S_CORR2G,
             engine.S_VSQG, engine.S_REGSQG]
    assert sorted(bases) == list(range(16, 16 * 9, 16)) and engine.SCAL_SLOTS >= max(bases) + 16
