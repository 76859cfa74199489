"""Bit-exact gates for the byte/integer ends (transform, tensor2im, seg aggregate + posneg mask)."""
import os

import numpy as np
import pytest
import torch

from oracle import pixel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import ops as _ops
    return _ops


def test_transform_and_tensor2im_bit_exact_vs_reference_golden(ops):
    z = np.load(os.path.join(GOLD, "pixel_ends.npz"))
    t = ops.u8_to_f32(torch.from_numpy(z["img"])[None].cuda()).cpu().numpy()
    assert np.array_equal(t, z["transform"])
    u8 = ops.f32_to_u8(torch.from_numpy(z["f"]).cuda()).cpu().numpy()[0]
    assert np.array_equal(u8, z["tensor2im"])


def test_all_u8_values_round_trip(ops):
    img = np.arange(256, dtype=np.uint8).repeat(3).reshape(1, 16, 16, 3)
    t = ops.u8_to_f32(torch.from_numpy(img).cuda())
    assert np.array_equal(t.cpu().numpy(), pixel.transform(img[0]))
    back = ops.f32_to_u8(t).cpu().numpy()
    assert np.array_equal(back[0], pixel.tensor2im(t.cpu().numpy()))


@pytest.mark.parametrize("N,H,W", [(1, 64, 64), (3, 512, 512)])
def test_seg_finish_bit_exact(ops, N, H, W):
    rng = np.random.default_rng(3)
    segs = [np.tanh(rng.standard_normal((N, 3, H, W)).astype(np.float32) * 2) for _ in range(5)]
    for weights in ([0.2] * 5, [0.5, 0.0, 0.0, 0.0, 0.5]):
        f32, u8, mask = ops.seg_finish([torch.from_numpy(s).cuda() for s in segs], weights, 120)
        ref = pixel.seg_aggregate(segs, weights)
        assert np.array_equal(f32.cpu().numpy(), ref)
        ref_u8 = pixel.tensor2im_batch(ref)
        assert np.array_equal(u8.cpu().numpy(), ref_u8)
        for i in range(N):
            assert np.array_equal(mask[i].cpu().numpy(), pixel.create_posneg_mask(ref_u8[i], 120))


def test_tile_gray_variance_matches_reference_golden(ops):
    """is_empty() statistic on the device vs values recorded from the reference's image_variance_gray."""
    z = np.load(os.path.join(GOLD, "tiler.npz"))
    v = ops.tile_gray_variance(torch.from_numpy(z["var_imgs"]).cuda())
    assert np.abs(v - z["var_vals"]).max() < 1e-6
    assert v[1] == 0.0 and (v[1] < 9) and (v[0] >= 9)      # constant tile is "empty", random tile is not


def test_tile_gray_variance_drops_saturated_pixels(ops):
    """Device statistic of is_empty() on tiles with luma 0 / 255 pixels (tests/golden/variance.npz, recorded from the
    reference): exact integer sums over the unsaturated pixels only."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "variance.npz"))
    v = ops.tile_gray_variance(torch.from_numpy(z["imgs"]).cuda())
    assert np.allclose(v, z["var"], rtol=1e-9, atol=1e-9)
    assert ((v < 9) == z["empty"]).all()
