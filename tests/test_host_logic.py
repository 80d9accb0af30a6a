"""Host-side planning logic that needs no GPU: tile / geometry cost models of fullbatchtraining_b200.ops."""
import pytest

from fullbatchtraining_b200 import ops


def test_pixel_tile_covers_128_pixels():
    for h, w in [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (64, 64)]:
        tw, th, tn = ops.pixel_tile(h, w)
        assert tw == w and tw * th * tn == 128 and (tn == 1 or th == h)
    with pytest.raises(RuntimeError):
        ops.pixel_tile(24, 24)


def test_choose_n_tile_follows_the_cost_model():
    # 64-channel stage: only 64 divides; 4x4 stage (16 M tiles): 64-wide stacked tiles fill more SMs than 128-wide ones
    assert ops.choose_n_tile(1024, 64, 2, 2) == 64
    assert ops.choose_n_tile(16, 512, 2, 2) == 64
    # 8x8 stage: 64 M tiles x 2 N tiles of 128 = one wave
    assert ops.choose_n_tile(64, 256, 2, 2) == 128
    for m, n in [(1, 64), (7, 192), (300, 2048)]:
        assert n % ops.choose_n_tile(m, n) == 0


def test_halo_geometry_is_opt_in_and_consistent(monkeypatch):
    monkeypatch.delenv("FB_HALO", raising=False)
    assert ops.halo_geometry(128, 32, 32, 3, 1, 64) is None
    monkeypatch.setenv("FB_HALO", "1")
    assert ops.halo_geometry(128, 32, 32, 3, 1, 64) == (1, 2, 64)
    assert ops.halo_geometry(128, 16, 16, 3, 1, 128) == (1, 2, 128)
    imgs, halves, nt = ops.halo_geometry(128, 8, 8, 3, 1, 256)
    assert imgs == 2 and 128 % (halves * imgs) == 0 and 256 % nt == 0
    assert ops.halo_geometry(128, 4, 4, 3, 1, 512)[0] == 8
    assert ops.halo_geometry(4, 4, 4, 3, 1, 512) is None       # fewer images than one interleaved half
    assert ops.halo_geometry(128, 32, 32, 1, 1, 64) is None    # 1x1
    assert ops.halo_geometry(128, 32, 32, 3, 2, 64) is None    # stride 2


def test_wgrad_halo_eligibility(monkeypatch):
    monkeypatch.delenv("FB_WGRAD_HALO", raising=False)
    assert ops.Conv2dPlan._wgrad_halo(3, 1, (32, 4, 1), 2)
    assert ops.Conv2dPlan._wgrad_halo(3, 1, (16, 8, 1), 2)
    assert not ops.Conv2dPlan._wgrad_halo(3, 1, (8, 8, 2), 2)   # tiles span two images
    assert not ops.Conv2dPlan._wgrad_halo(1, 1, (32, 4, 1), 2)
    assert not ops.Conv2dPlan._wgrad_halo(3, 2, (16, 8, 1), 2)
    monkeypatch.setenv("FB_WGRAD_HALO", "0")
    assert not ops.Conv2dPlan._wgrad_halo(3, 1, (32, 4, 1), 2)


def test_partial_workspace_bound_covers_every_mode():
    for case in [(128, 32, 32, 64, 64, 3, 1), (128, 16, 16, 128, 128, 3, 1), (128, 8, 8, 256, 256, 3, 1),
                 (128, 32, 32, 64, 128, 3, 2), (32, 8, 8, 1024, 256, 1, 1)]:
        n, h, w, cin, cout, k, stride = case
        assert ops.Conv2dPlan.partial_elems(*case) >= cout * k * k * cin
