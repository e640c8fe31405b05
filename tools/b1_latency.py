#!/usr/bin/env python
"""B = 1 latency of the real call site (run on the GPU box):  plsvo::SparseImgAlign(4, 2, 30, GaussNewton, false, false)
.run(ref_frame, cur_frame) exactly as src/frame_handler_mono.cpp:272-274 makes it — reference-typed Frame / Feature objects
through the signature-preserving shim (oracle/shimref_harness.cpp) onto the C ABI with batch = 1 — next to the same call
answered by the reference's own sparse_img_align.cpp on one host thread.

    python tools/b1_latency.py [pairs] > profiles/r02_b1_latency.txt
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import bench  # noqa: E402
import plsvo_b200  # noqa: E402,F401
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402


def pct(x):
    return "p50 %.3f ms, p90 %.3f ms, p99 %.3f ms, mean %.3f ms" % (*(1e3 * np.percentile(x, [50, 90, 99])), 1e3 * np.mean(x))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    d = synth.make_align_batch(batch=n, n_pts=300, n_segs=80, device="cuda", seed=5100)
    lib = oracle_lib.load_shimref(abi)
    lib.plsvo_shimref_run_seconds.restype = C.c_int
    lib.plsvo_shimref_run_seconds.argtypes = [C.POINTER(C.c_double), C.c_int]
    oracle_lib.shimref_align(abi, bench.subset(d, 8))  # warm: context, module load, first allocations
    got = oracle_lib.shimref_align(abi, d)
    secs = np.zeros(n)
    m = lib.plsvo_shimref_run_seconds(secs.ctypes.data_as(C.POINTER(C.c_double)), n)
    assert m == n
    cpu = []
    fn = oracle_lib.ref_align if oracle_lib.ref_available() else oracle_lib.align
    ref_out = np.zeros((n, 7))
    for b in range(n):
        one = bench.subset(d, b + 1)
        one = _last(one)
        t0 = time.perf_counter()
        r = fn(abi, one, n_threads=1)
        cpu.append(time.perf_counter() - t0)
        ref_out[b] = r.T_cur_w[0]
    cpu = np.array(cpu)
    ang, rel = synth.pose_error(got.T_cur_w, ref_out)
    print("workload: VGA, levels 4->2, 300 points + 80 segments, one frame pair per call (B = 1), %d calls" % n)
    print("B200, plsvo::SparseImgAlign::run via the shim (host objects -> pack -> H2D -> kernel on one CTA -> D2H -> write back): " + pct(secs))
    print("CPU, the reference's sparse_img_align.cpp (%s), one thread:                                                       " % (
        "oracle/_ref" if oracle_lib.ref_available() else "oracle port") + pct(cpu))
    print("speed-up at B = 1 (p50): %.1fx; parity: %d / %d pairs inside 1e-5 rad / 1e-4 rel-t, n_tracked equal: %s" % (
        np.percentile(cpu, 50) / np.percentile(secs, 50), int(((ang <= 1e-5) & (rel <= 1e-4)).sum()), n, "n/a"))
    pose_optimizer(n, lib)


def pose_optimizer(n, lib):
    """The second call of the per-frame path, pose_optimizer::optimizeGaussNewton(2.0, 10, false, frame, ...)
    (src/frame_handler_mono.cpp:327-329), one frame per call."""
    pd = synth.make_poseopt_batch(batch=n, n_pts=300, n_segs=80, seed=5200)
    pp = abi.poseopt_params(2.0, 10, -1)
    lib.plsvo_shimref_poseopt_seconds.restype = C.c_int
    lib.plsvo_shimref_poseopt_seconds.argtypes = [C.POINTER(C.c_double), C.c_int]
    oracle_lib.shimref_poseopt(abi, pd, pp)  # warm
    got = oracle_lib.shimref_poseopt(abi, pd, pp)
    secs = np.zeros(n)
    assert lib.plsvo_shimref_poseopt_seconds(secs.ctypes.data_as(C.POINTER(C.c_double)), n) == n
    fn = oracle_lib.ref_poseopt if oracle_lib.ref_available() else oracle_lib.poseopt
    cpu, ref_T = [], np.zeros((n, 7))
    import copy

    for b in range(n):
        one = copy.copy(pd)
        for name, v in vars(pd).items():
            if isinstance(v, np.ndarray) and v.shape[:1] == (n,):
                setattr(one, name, np.ascontiguousarray(v[b:b + 1]))
        t0 = time.perf_counter()
        r = fn(abi, one, pp, n_threads=1)
        cpu.append(time.perf_counter() - t0)
        ref_T[b] = r.T_f_w[0]
    cpu = np.array(cpu)
    ang, rel = synth.pose_error(got.T_f_w, ref_T)
    print("\nworkload: 300 point + 80 line observations, <= 10 Gauss-Newton iterations, one frame per call (B = 1), %d calls" % n)
    print("B200, plsvo::pose_optimizer::optimizeGaussNewton via the shim: " + pct(secs))
    print("CPU, the reference's pose_optimizer.cpp, one thread:           " + pct(cpu))
    print("speed-up at B = 1 (p50): %.1fx; parity: %d / %d frames inside 1e-5 rad / 1e-4 rel-t" % (
        np.percentile(cpu, 50) / np.percentile(secs, 50), int(((ang <= 1e-5) & (rel <= 1e-4)).sum()), n))


def _last(sub):
    """the last pair of an AlignData subset as a batch of one"""
    import copy

    one = copy.copy(sub)
    for name in ("T_ref_w", "T_cur_w", "T_cur_w_gt", "pt_px", "pt_f", "pt_pos", "seg_spx", "seg_epx", "seg_sf", "seg_ef", "seg_spos",
                 "seg_epos", "seg_length"):
        setattr(one, name, np.ascontiguousarray(getattr(sub, name)[-1:]))
    one.ref_pyr = {l: np.ascontiguousarray(v[-1:]) for l, v in sub.ref_pyr.items()}
    one.cur_pyr = {l: np.ascontiguousarray(v[-1:]) for l, v in sub.cur_pyr.items()}
    return one


if __name__ == "__main__":
    main()
