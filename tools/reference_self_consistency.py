#!/usr/bin/env python
"""How reproducible is the reference itself?  (authoring container only: needs /root/reference)

Builds the reference's own translation units (oracle/ref_harness.cpp, oracle/refdeps) twice:
  strict : -O3 -ffp-contract=off            (oracle/_ref/libplsvo_ref.so, the parity anchor)
  release: the reference's own release flags, CMakeLists.txt:25,36 — -O3 -march=native
           -fno-signed-zeros -fno-math-errno -funroll-loops (GCC contracts a*b+c into FMA)
and counts, over seeded C2-shaped pairs, how many final poses differ by more than the parity
tolerance (1e-5 rad / 1e-4 relative translation) and how many pairs run a different number of
Gauss-Newton iterations.  The termination test `new_chi2 > chi2_` compares two float sums that
often agree to ~1e-6, so any change of rounding flips it for a fraction of a percent of pairs.

    python tools/reference_self_consistency.py [n_batches] > profiles/r01_reference_self_consistency.txt
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import plsvo_b200  # noqa: E402,F401
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402

REF = os.environ.get("PLSVO_REFERENCE_ROOT", "/root/reference")
NATIVE = os.path.join(ROOT, "oracle", "_ref", "libplsvo_ref_release_flags.so")
FLAGS = "-O3 -march=native -fno-signed-zeros -fno-math-errno -funroll-loops -fomit-frame-pointer -std=c++17 -fPIC -w -DNDEBUG"


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    if not oracle_lib.build_ref():
        raise SystemExit("needs /root/reference")
    srcs = [os.path.join(REF, "src", f) for f in ("sparse_img_align.cpp", "pose_optimizer.cpp", "feature.cpp", "feature_alignment.cpp", "matcher.cpp", "config.cpp", "feature3D_impl.cpp", "depth_filter.cpp")]
    subprocess.check_call(["g++", *FLAGS.split(), "-I" + os.path.join(ROOT, "oracle", "refdeps"), "-I" + os.path.join(REF, "include"),
                           "-shared", "-o", NATIVE, *srcs, os.path.join(ROOT, "oracle", "ref_harness.cpp"), "-lpthread"])
    nat = C.CDLL(NATIVE)
    P = C.POINTER
    nat.plsvo_ref_align_batch.restype = C.c_int
    nat.plsvo_ref_align_batch.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), P(abi.AlignResult), C.c_int]
    print("strict  :", oracle_lib.REF_LIB_PATH, "(-O3 -ffp-contract=off)")
    print("release :", NATIVE, f"({FLAGS})")
    tot = bad = itd = 0
    threads = os.cpu_count() or 1
    for seed in range(4000, 4000 + n_batches):
        d = synth.make_align_batch(cam=synth.VGA, batch=256, n_pts=300, n_segs=80, seed=seed)
        strict = oracle_lib.ref_align(abi, d, n_threads=threads)
        params = abi.align_params(d.max_level, d.min_level)
        batch, keep = abi.make_align_batch(d)
        out = abi.AlignOut(d.batch, d.n_segs)
        assert nat.plsvo_ref_align_batch(C.byref(batch), C.byref(params), C.byref(out.struct), threads) == 0
        ang, rel = synth.pose_error(torch.tensor(out.T_cur_w), torch.tensor(strict.T_cur_w))
        ang, rel = np.asarray(ang), np.asarray(rel)
        ok = (ang <= 1e-5) & (rel <= 1e-4)
        nd = int((out.iters != strict.iters).any(axis=1).sum())
        tot += len(ok)
        bad += int((~ok).sum())
        itd += nd
        print(f"seed {seed}: 256 pairs, out of tolerance {int((~ok).sum())}, iteration counts differ {nd}, "
              f"max rot {ang.max():.3e} rad, max rel t {rel.max():.3e}", flush=True)
    print(f"TOTAL {tot} pairs: {bad} out of tolerance ({100.0 * bad / tot:.2f} %) between two builds of the reference's own code; "
          f"{itd} pairs ({100.0 * itd / tot:.2f} %) run a different number of GN iterations")


if __name__ == "__main__":
    main()
