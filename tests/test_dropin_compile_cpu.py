"""Drop-in proof at the source level (authoring container only: needs /root/reference).

The reference's caller of both hot-path symbols, src/frame_handler_mono.cpp, is compiled UNMODIFIED with
pl-svo_b200/host/overlay/ in front of the reference's include path — i.e. with plsvo/sparse_img_align.h and
plsvo/pose_optimizer.h replaced by the B200 shim, exactly what INTEGRATION.md §2 tells a maintainer to do — and the shim
itself is compiled in -DPLSVO_SHIM_WITH_REFERENCE_HEADERS mode against the reference's own Frame / Feature / SE3 types.
Every SparseImgAlign / pose_optimizer symbol the caller leaves undefined must then be defined by the shim object
(same mangled names = same signatures).  Third-party headers come from oracle/refdeps (stand-ins)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PLSVO_REFERENCE_ROOT", "/root/reference")
HOST = os.path.join(ROOT, "pl-svo_b200", "host")
COMMON = ["g++", "-std=c++17", "-O0", "-fPIC", "-w", "-DPLSVO_SHIM_WITH_REFERENCE_HEADERS",
          "-I", os.path.join(HOST, "overlay"), "-I", os.path.join(ROOT, "oracle", "refdeps"), "-I", os.path.join(REF, "include"),
          "-I", os.path.join(ROOT, "include"), "-I", HOST]


def _syms(obj, flag):
    out = subprocess.check_output(["nm", flag, obj]).decode().splitlines()
    return {line.split()[-1] for line in out if line.strip()}


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "frame_handler_mono.cpp")), reason="/root/reference is absent")
def test_reference_caller_compiles_and_links_against_the_shim(tmp_path):
    caller = str(tmp_path / "frame_handler_mono.o")
    shim = str(tmp_path / "plsvo_shim.o")
    subprocess.check_call(COMMON + ["-c", os.path.join(REF, "src", "frame_handler_mono.cpp"), "-o", caller])
    subprocess.check_call(COMMON + ["-c", os.path.join(HOST, "plsvo_shim.cpp"), "-o", shim])
    wanted = {s for s in _syms(caller, "-u") if "SparseImgAlign" in s or "pose_optimizer" in s}
    provided = _syms(shim, "--defined-only")
    assert len(wanted) >= 3, wanted  # constructor, run, optimizeGaussNewton (9-argument overload)
    assert wanted <= provided, sorted(wanted - provided)
    demangled = subprocess.check_output(["c++filt"] + sorted(wanted)).decode()
    assert "plsvo::SparseImgAlign::run(" in demangled and "plsvo::pose_optimizer::optimizeGaussNewton(" in demangled
