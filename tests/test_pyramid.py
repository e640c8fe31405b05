"""Image pyramid (frame_utils::createImgPyramid = repeated vk::halfSample): byte work, bit-exact."""
import numpy as np
import pytest
import torch


def test_oracle_half_sample_matches_integer_definition(oracle, abi, synth):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (3, 61, 83), dtype=np.uint8)
    lv = oracle.pyramid(abi, img, 4)
    ref = synth.build_pyramid(torch.from_numpy(img), 4)
    for a, b in zip(lv, ref):
        np.testing.assert_array_equal(a, b.numpy())
    # truncation, not rounding: block (0,0,1,2) -> 0
    tiny = np.array([[[0, 0], [1, 2]]], np.uint8)
    assert oracle.pyramid(abi, tiny, 2)[1][0, 0, 0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_levels", [((4, 480, 640), 5), ((2, 720, 1280), 6), ((3, 479, 641), 5), ((1, 70, 130), 7)])
def test_gpu_pyramid_is_bit_exact(pkg, abi, oracle, shape, n_levels):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    gpu = pkg.createImgPyramid(img, n_levels)
    ref = oracle.pyramid(abi, img, n_levels)
    assert len(gpu) == n_levels
    for l in range(n_levels):
        assert gpu[l].shape == ref[l].shape
        np.testing.assert_array_equal(gpu[l], ref[l])


@pytest.mark.gpu
def test_gpu_pyramid_feeds_alignment(pkg, abi, synth, oracle, gen_device):
    """Pyramids built on the device are the ones the generator builds on the host."""
    d = synth.make_align_batch(batch=2, n_pts=50, n_segs=8, device=gen_device, seed=4100, keep_levels_only=False)
    lv = pkg.createImgPyramid(d.cur_pyr[0], 5)
    for l in range(5):
        np.testing.assert_array_equal(lv[l], d.cur_pyr[l])
