"""Error behaviour of the C ABI on a live device: misuse returns PLSVO_ERR_* with a message, never crashes,
and the reference's own "nothing to do" conventions are status bits, not errors."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_state_and_argument_errors(pkg, abi, synth, gen_device):
    ctx = pkg.Context(0)
    lib = ctx.lib
    params = abi.align_params(4, 2, 30)
    # launch / download before any upload
    assert lib.plsvo_align_launch(ctx.handle, C.byref(params)) == abi.ERR_STATE
    assert b"before plsvo_align_upload" in lib.plsvo_last_error(ctx.handle)
    out = abi.AlignOut(1, 0)
    assert lib.plsvo_align_download(ctx.handle, C.byref(out.struct)) == abi.ERR_STATE
    pp = abi.poseopt_params()
    assert lib.plsvo_poseopt_launch(ctx.handle, C.byref(pp)) == abi.ERR_STATE
    # NULL arguments
    assert lib.plsvo_align_upload(ctx.handle, None) == abi.ERR_INVALID
    assert lib.plsvo_sync(None) == abi.ERR_INVALID
    # inconsistent batch descriptions
    d = synth.make_align_batch(cam=synth.QVGA, batch=2, n_pts=16, n_segs=4, max_level=3, min_level=1, margin=32,
                               device=gen_device, seed=1)
    batch, keep = abi.make_align_batch(d)
    batch.batch = 0
    assert lib.plsvo_align_upload(ctx.handle, C.byref(batch)) == abi.ERR_INVALID
    batch.batch = 2
    saved = C.cast(batch.pt_px, C.c_void_p).value  # the field object aliases the struct: keep the address
    batch.pt_px = None
    assert lib.plsvo_align_upload(ctx.handle, C.byref(batch)) == abi.ERR_INVALID
    assert b"point arrays" in lib.plsvo_last_error(ctx.handle)
    batch.pt_px = C.cast(C.c_void_p(saved), C.POINTER(C.c_double))
    assert lib.plsvo_align_upload(ctx.handle, C.byref(batch)) == abi.OK
    # level range that was neither uploaded nor derivable (level 0 has nothing below it to be derived from; levels above
    # the uploaded ones would simply be derived on the device) / nonsense parameters
    bad = abi.align_params(3, 0, 30)
    assert lib.plsvo_align_launch(ctx.handle, C.byref(bad)) == abi.ERR_INVALID
    assert b"no lower level" in lib.plsvo_last_error(ctx.handle)
    bad = abi.align_params(1, 3, 30)
    assert lib.plsvo_align_launch(ctx.handle, C.byref(bad)) == abi.ERR_INVALID
    bad = abi.align_params(3, 1, 0)
    assert lib.plsvo_align_launch(ctx.handle, C.byref(bad)) == abi.ERR_INVALID
    # and the context is still usable afterwards
    good = abi.align_params(3, 1, 30)
    assert lib.plsvo_align_launch(ctx.handle, C.byref(good)) == abi.OK
    out = abi.AlignOut(2, 4)
    assert lib.plsvo_align_download(ctx.handle, C.byref(out.struct)) == abi.OK
    assert (out.status == 0).all() and (out.n_tracked > 0).all()
    # device ordinal out of range
    h = C.c_void_p()
    assert lib.plsvo_ctx_create(99, None, C.byref(h)) == abi.ERR_INVALID
    ctx.close()


def test_pinned_host_alloc_roundtrip(abi):
    lib = abi.load_library()
    p = C.c_void_p()
    assert lib.plsvo_host_alloc(C.byref(p), 1 << 20) == abi.OK and p.value
    C.memset(p, 7, 1 << 20)
    assert lib.plsvo_host_free(p) == abi.OK
