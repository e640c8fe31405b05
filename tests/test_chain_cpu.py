"""Frame chains on the CPU side: the generator, the one-stack view and the ABI struct (no device needed).

A chain batch replays one camera sequence — pair b = (frame b, frame b+1), src/frame_handler_mono.cpp:176,272 — and the
C ABI accepts it as ONE stack of B+1 frames (PLSVO_ALIGN_FRAME_CHAIN, include/plsvo_b200.h)."""
import ctypes as C
import re
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain_batch_is_a_chain_and_the_one_stack_view_holds_every_frame_once(synth):
    data = synth.make_chain_batch(batch=5, n_pts=40, n_segs=8, seed=5300)
    for l in data.ref_pyr:
        np.testing.assert_array_equal(data.cur_pyr[l][:-1], data.ref_pyr[l][1:])
    np.testing.assert_array_equal(data.T_cur_w_gt[:-1], data.T_ref_w[1:])
    np.testing.assert_array_equal(data.T_cur_w, data.T_ref_w)  # initial guess = previous pose (frame_handler_mono.cpp:266)
    frames = synth.chain_frames(data)
    for l, f in frames.items():
        assert f.shape[0] == data.batch + 1 and f.flags["C_CONTIGUOUS"]
        np.testing.assert_array_equal(f[:-1], data.ref_pyr[l])
        np.testing.assert_array_equal(f[1:], data.cur_pyr[l])
    assert list(synth.chain_frames(data, levels=[2])) == [2]
    # consecutive frames differ (the camera moves) but overlap (small steps)
    d = np.abs(frames[2][1:].astype(int) - frames[2][:-1].astype(int)).mean()
    assert 0.5 < d < 60


def test_independent_pairs_are_not_a_chain(synth):
    data = synth.make_align_batch(batch=3, n_pts=10, n_segs=2, seed=5301)
    with pytest.raises(ValueError):
        synth.chain_frames(data)


def test_abi_struct_for_a_chain(abi, synth):
    import copy

    data = synth.make_chain_batch(batch=4, n_pts=20, n_segs=4, seed=5302)
    two, _k2 = abi.make_align_batch(data)
    assert two.flags == 0 and two.ref_img[2] and two.cur_img[2]
    one_data = copy.copy(data)
    one_data.frame_pyr = synth.chain_frames(data, levels=[2])
    one, _k1 = abi.make_align_batch(one_data)
    assert one.flags == abi.ALIGN_FRAME_CHAIN
    assert not one.cur_img[2] and not one.ref_img[3] and not one.ref_img[4]
    f = one_data.frame_pyr[2]
    assert C.cast(one.ref_img[2], C.c_void_p).value == f.ctypes.data
    assert one.img_pitch[2] == f.strides[1] and one.img_stride[2] == f.strides[0]
    # the flag occupies what used to be a reserved int: the layout of every other field is unchanged
    assert abi.AlignBatch.flags.offset == 12 and abi.AlignBatch.cam.offset == 16


def test_header_and_python_mirror_agree_on_the_flag(abi):
    text = open(os.path.join(ROOT, "include", "plsvo_b200.h")).read()
    m = re.search(r"#define\s+PLSVO_ALIGN_FRAME_CHAIN\s+(\d+)", text)
    assert m and int(m.group(1)) == abi.ALIGN_FRAME_CHAIN
    assert re.search(r"int32_t\s+flags;", text)


def test_oracle_result_on_a_chain_does_not_depend_on_how_the_frames_are_stored(abi, synth, oracle):
    """The oracle reads the two stacks; pointing its cur stack into the one-stack array (frame b+1) is the same input."""
    import copy

    data = synth.make_chain_batch(batch=3, n_pts=60, n_segs=10, seed=5303)
    a = oracle.align(abi, data, abi.align_params(4, 2, 30), n_threads=2)
    frames = synth.chain_frames(data)
    alias = copy.copy(data)
    alias.ref_pyr = {l: f[:-1] for l, f in frames.items()}
    alias.cur_pyr = {l: f[1:] for l, f in frames.items()}
    b = oracle.align(abi, alias, abi.align_params(4, 2, 30), n_threads=2)
    np.testing.assert_array_equal(a.T_cur_w, b.T_cur_w)
    np.testing.assert_array_equal(a.iters, b.iters)
    assert (a.n_tracked > 0).all()
    ang, _ = synth.pose_error(a.T_cur_w, data.T_cur_w_gt)
    assert np.median(ang) < 2e-3
