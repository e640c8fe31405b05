#!/usr/bin/env python
"""Measurement of the SURVEY §8f ("next") kernels on one B200: pyramid construction, align2D, align1D.

For each kernel: device time of the kernel alone (CUDA events on the context's stream, plsvo_last_kernel_ms;
median of `--reps` calls after warm-up), the end-to-end time of the host-in/host-out C-ABI call, the algorithmic
bytes the kernel must move (stated below) against the measured HBM copy peak, bit-exact parity against the CPU
oracle on the timed inputs, and the CPU oracle's rate on all host threads beside it.  Prints one JSON object.

    python tools/bench_next.py > profiles/r01_next_kernels.json

Algorithmic bytes
  pyramid : level 0 read once + levels 1..n-1 written once  = W*H*(1 + 1/4 + ... ) bytes per image
  align2D : per feature 100 + 64 (patches) + 16 (px in) + 8 + 17 (level/index, outputs) = 205 B, plus the
            9x9 u8 image footprint per executed iteration (81 B x iterations, counted from the oracle's run)
  align1D : as align2D + 8 B direction + 8 B h_inv
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import plsvo_b200  # noqa: E402
from plsvo_b200 import abi, api, synth  # noqa: E402
import oracle_lib  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def best_threads(fn, n_items):
    hw = max(1, oracle_lib.load(abi).plsvo_oracle_hardware_threads())
    best = None
    for th in sorted({hw, max(1, hw // 2), max(1, hw // 4), max(1, hw // 8)}, reverse=True):
        fn(th)
        t0 = time.perf_counter()
        fn(th)
        r = n_items / (time.perf_counter() - t0)
        if best is None or r > best[0]:
            best = (r, th)
    return best


def timed(ctx, call, reps):
    call()
    call()
    k, e = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        e.append((time.perf_counter() - t0) * 1e3)
        k.append(ctx.last_kernel_ms())
    return float(np.median(k)), float(np.median(e))


def feature_case(n, seed, dev):
    rng = np.random.default_rng(seed)
    cam = synth.VGA
    n_img = 16
    poses = synth.pose7_from_Rt(*synth.se3_exp_Rt(torch.tensor(rng.uniform(-0.05, 0.05, (n_img, 6)), dtype=torch.float64)))
    img0 = synth.Scene().render(cam, poses.to(dev)).cpu()
    pyr = {l: np.ascontiguousarray(p.numpy()) for l, p in enumerate(synth.build_pyramid(img0, 3))}
    idx = rng.integers(0, n_img, n).astype(np.int32)
    lvl = rng.integers(0, 3, n).astype(np.int32)
    xi = np.array([rng.integers(12, (cam.width >> l) - 12) for l in lvl])
    yi = np.array([rng.integers(12, (cam.height >> l) - 12) for l in lvl])
    border = np.zeros((n, 10, 10), np.uint8)
    for l in range(3):
        sel = np.nonzero(lvl == l)[0]
        for dy in range(10):
            for dx in range(10):
                border[sel, dy, dx] = pyr[l][idx[sel], yi[sel] - 5 + dy, xi[sel] - 5 + dx]
    ref = np.ascontiguousarray(border[:, 1:9, 1:9])
    px0 = np.stack([xi, yi], -1) + rng.uniform(-1.5, 1.5, (n, 2))
    ang = rng.uniform(0, 2 * np.pi, n)
    dirs = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)
    return cam, pyr, idx, lvl, border, ref, px0, dirs


_PINNED = []


def pin(a):
    """Page-locked copy of a numpy array (kept alive for the life of the process): the host-in/host-out calls below are
    timed from pinned memory, as an integrating caller that cares about the copy would hold its frames."""
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    _PINNED.append(t)
    return t.numpy()


def pin_obj(d):
    """Every numpy array attribute (and dict-of-arrays attribute) of a synth data object, replaced by a pinned copy."""
    for k, v in list(vars(d).items()):
        if isinstance(v, np.ndarray) and v.size:
            setattr(d, k, pin(v))
        elif isinstance(v, dict) and v and all(isinstance(x, np.ndarray) for x in v.values()):
            setattr(d, k, {kk: pin(x) for kk, x in v.items()})
    return d


def measure(reps=10, images=256, features=400000, struct_points=200000, ctx=None, dev=None):
    """All §8f kernels once; returns the result dict (see the module docstring)."""
    import types

    args = types.SimpleNamespace(reps=reps, images=images, features=features)
    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = dev or torch.device("cuda", 0)
    ctx = ctx or api.default_context()
    olib = oracle_lib.load(abi)
    olib.plsvo_oracle_pyramid_batch.restype = C.c_int
    olib.plsvo_oracle_align2d_batch.restype = C.c_int
    olib.plsvo_oracle_align1d_batch.restype = C.c_int
    peak, peak_src = hbm_peak()
    res = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "hbm_peak_source": peak_src, "reps": args.reps}

    # ---- pyramid -------------------------------------------------------------------------------------
    cam = synth.VGA
    B, L = args.images, 5
    rng = np.random.default_rng(1)
    img0 = pin(rng.integers(0, 256, (B, cam.height, cam.width), dtype=np.uint8))
    levels = {}

    def run_pyr():
        levels["gpu"] = api.createImgPyramid(img0, L, ctx)

    k_ms, e_ms = timed(ctx, run_pyr, args.reps)
    ref_levels = oracle_lib.pyramid(abi, img0[:8], L)
    exact = all(np.array_equal(levels["gpu"][l][:8], ref_levels[l]) for l in range(1, L))
    bytes_img = sum((cam.width >> l) * (cam.height >> l) for l in range(L))
    pb = abi.PyramidBatch(B, cam.width, cam.height, L, img0.ctypes.data_as(C.POINTER(C.c_uint8)), img0.strides[1], img0.strides[0])
    outs = [None] + [np.empty((B, cam.height >> l, cam.width >> l), np.uint8) for l in range(1, L)]
    pr = abi.PyramidResult()
    for l in range(1, L):
        pr.level[l] = outs[l].ctypes.data_as(C.POINTER(C.c_uint8))
        pr.pitch[l] = outs[l].strides[1]
        pr.stride[l] = outs[l].strides[0]
    rate, th = best_threads(lambda t: olib.plsvo_oracle_pyramid_batch(C.byref(pb), C.byref(pr), t), B)
    res["pyramid"] = {
        "workload": f"{B} VGA images, {L} levels (frame_utils::createImgPyramid)", "kernel_ms": k_ms, "e2e_ms": e_ms,
        "images_per_s_kernel": B / (k_ms * 1e-3), "images_per_s_e2e": B / (e_ms * 1e-3),
        "algorithmic_bytes": B * bytes_img, "achieved_gbs": B * bytes_img / (k_ms * 1e-3) / 1e9,
        "roofline_frac": B * bytes_img / (k_ms * 1e-3) / 1e9 / peak, "bit_exact_vs_oracle": bool(exact),
        "cpu_oracle_images_per_s": rate, "cpu_threads": th}

    # ---- align2D / align1D -----------------------------------------------------------------------------
    n = args.features
    cam, pyr, idx, lvl, border, ref, px0, dirs = feature_case(n, 2, dev)
    pyr = {l: pin(v) for l, v in pyr.items()}
    idx, lvl, border, ref, px0, dirs = pin(idx), pin(lvl), pin(border), pin(ref), pin(px0), pin(dirs)
    feats, keep = abi.make_align2d_batch(pyr, idx, lvl, border, ref, np.ascontiguousarray(px0), 10, cam.width, cam.height)
    o_px = np.zeros((n, 2))
    o_cv = np.zeros(n, np.uint8)
    o_h = np.zeros(n)
    f64p, u8p = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    r2 = abi.Align2DResult(o_px.ctypes.data_as(f64p), o_cv.ctypes.data_as(u8p))
    b1 = abi.Align1DBatch(feats, dirs.ctypes.data_as(C.POINTER(C.c_float)))
    r1 = abi.Align1DResult(o_px.ctypes.data_as(f64p), o_cv.ctypes.data_as(u8p), o_h.ctypes.data_as(f64p))
    out = {}

    def run2():
        out["a2"] = api.feature_alignment.align2D(pyr, idx, lvl, border, ref, 10, px0, cam.width, cam.height, ctx)

    def run1():
        out["a1"] = api.feature_alignment.align1D(pyr, idx, lvl, dirs, border, ref, 10, px0, cam.width, cam.height, ctx)

    for name, run, cpu_call, extra in (
            ("align2d", run2, lambda t: olib.plsvo_oracle_align2d_batch(C.byref(feats), C.byref(r2), t), 0),
            ("align1d", run1, lambda t: olib.plsvo_oracle_align1d_batch(C.byref(b1), C.byref(r1), t), 16)):
        k_ms, e_ms = timed(ctx, run, args.reps)
        rate, th = best_threads(cpu_call, n)  # leaves the oracle's outputs in o_px / o_cv (/ o_h)
        got = out["a2" if name == "align2d" else "a1"]
        fin = np.isfinite(o_px).all(axis=1)
        exact = np.array_equal(got[0], o_cv.astype(bool)) and np.array_equal(got[1][fin], o_px[fin])
        if name == "align1d":
            exact = exact and np.array_equal(got[2], o_h)
        # iterations are not an output of the reference function: bound the image traffic with n_iter passes
        alg = n * (205 + extra) + n * 81 * 10
        res[name] = {
            "workload": f"{n} features over 16 VGA frames, levels 0-2, n_iter 10 (feature_alignment::{name.replace('d', 'D')})",
            "kernel_ms": k_ms, "e2e_ms": e_ms, "features_per_s_kernel": n / (k_ms * 1e-3), "features_per_s_e2e": n / (e_ms * 1e-3),
            "algorithmic_bytes_upper_bound": alg, "achieved_gbs_upper_bound": alg / (k_ms * 1e-3) / 1e9,
            "roofline_frac_upper_bound": alg / (k_ms * 1e-3) / 1e9 / peak, "bit_exact_vs_oracle": bool(exact),
            "converged_frac": float(got[0].mean()), "cpu_oracle_features_per_s": rate, "cpu_threads": th}
    # ---- Matcher::findMatchDirect (affine warp + patch + align) --------------------------------------------
    m = args.features // 2
    md = pin_obj(synth.make_match_batch(n=m, n_ref=8, n_cur=8, n_pyr_levels=3, seed=3, device=dev))
    matcher = api.Matcher(10, ctx)
    mo = {}

    def run_m():
        mo["gpu"] = matcher.findMatchDirect(md)

    k_ms, e_ms = timed(ctx, run_m, args.reps)
    olib.plsvo_oracle_match_direct_batch.restype = C.c_int
    mb, keep_m = abi.make_match_batch(md)
    cpu_out = abi.MatchOut(m)
    rate, th = best_threads(lambda t: olib.plsvo_oracle_match_direct_batch(C.byref(mb), C.byref(cpu_out.struct), t), m)
    fin = np.isfinite(cpu_out.px_cur).all(axis=1)
    exact = (np.array_equal(mo["gpu"].success, cpu_out.success) and np.array_equal(mo["gpu"].search_level, cpu_out.search_level)
             and np.array_equal(mo["gpu"].px_cur[fin], cpu_out.px_cur[fin]))
    # per candidate: inputs 2 idx + level + flag (13) + px,f,grad,pos,px_cur (96) + outputs (21) = 130 B; the warp reads
    # 100 x 4 reference bytes (bilinear taps, upper bound), align reads 81 B x iterations (upper bound n_iter)
    alg = m * (130 + 400 + 81 * 10)
    res["find_match_direct"] = {
        "workload": f"{m} candidates, 8 keyframes -> 8 current VGA frames, 3 pyramid levels, 25 % edgelets (Matcher::findMatchDirect)",
        "kernel_ms": k_ms, "e2e_ms": e_ms, "candidates_per_s_kernel": m / (k_ms * 1e-3), "candidates_per_s_e2e": m / (e_ms * 1e-3),
        "algorithmic_bytes_upper_bound": alg, "achieved_gbs_upper_bound": alg / (k_ms * 1e-3) / 1e9,
        "roofline_frac_upper_bound": alg / (k_ms * 1e-3) / 1e9 / peak, "bit_exact_vs_oracle": bool(exact),
        "success_frac": float(cpu_out.success.mean()), "cpu_oracle_candidates_per_s": rate, "cpu_threads": th}
    # ---- structure optimisation (Point::optimize / LineSeg::optimize) ------------------------------------
    sd = pin_obj(synth.make_structopt_batch(n_points=struct_points, n_segs=struct_points // 4, n_frames=32, seed=4))
    so = {}

    def run_s():
        so["gpu"] = api.optimizeStructure(sd, ctx)

    k_ms, e_ms = timed(ctx, run_s, args.reps)
    olib.plsvo_oracle_structopt_batch.restype = C.c_int
    sb, keep_s = abi.make_structopt_batch(sd)
    cpu_s = abi.StructOptOut(sb.n_points, sb.n_segs)
    nfeat = sb.n_points + sb.n_segs
    rate, th = best_threads(lambda t: olib.plsvo_oracle_structopt_batch(C.byref(sb), C.byref(cpu_s.struct), t), nfeat)
    exact = (np.array_equal(so["gpu"].pt_pos, cpu_s.pt_pos) and np.array_equal(so["gpu"].seg_spos, cpu_s.seg_spos)
             and np.array_equal(so["gpu"].seg_epos, cpu_s.seg_epos) and np.array_equal(so["gpu"].pt_iters, cpu_s.pt_iters))
    n_pt_obs, n_seg_obs = int(sd.pt_obs_begin[-1]), int(sd.seg_obs_begin[-1])
    # per executed iteration an observation costs its frame index + bearing(s) + the 56-byte pose (re-read, L2-resident)
    passes_pt = float(cpu_s.pt_iters.mean()), float(cpu_s.seg_iters.mean())
    alg = (n_pt_obs * (4 + 24 + 56) * passes_pt[0] + n_seg_obs * (4 + 48 + 56) * passes_pt[1]
           + sb.n_points * (24 + 24 + 8) + sb.n_segs * (48 + 48 + 8))
    res["structure_optimisation"] = {
        "workload": f"{sb.n_points} points ({n_pt_obs} observations) + {sb.n_segs} segments ({n_seg_obs} observations), 32 keyframes, "
                    "5 GN iterations (Point::optimize / LineSeg::optimize)",
        "kernel_ms": k_ms, "e2e_ms": e_ms, "features_per_s_kernel": nfeat / (k_ms * 1e-3), "features_per_s_e2e": nfeat / (e_ms * 1e-3),
        "algorithmic_bytes": alg, "achieved_gbs": alg / (k_ms * 1e-3) / 1e9, "roofline_frac": alg / (k_ms * 1e-3) / 1e9 / peak,
        "bit_exact_vs_oracle": bool(exact), "mean_iterations": passes_pt, "cpu_oracle_features_per_s": rate, "cpu_threads": th}
    # ---- depth-filter point-seed update (epipolar ZMSSD search + Bayesian update) ------------------------------
    ns = args.features // 4
    sdd = pin_obj(synth.make_seed_batch(n=ns, n_ref=8, n_cur=8, n_pyr_levels=3, seed=5, device=dev))
    dfilt = api.DepthFilter(ctx)
    dfo = {}

    def run_df():
        dfo["gpu"] = dfilt.updatePointSeeds(sdd)

    k_ms, e_ms = timed(ctx, run_df, args.reps)
    olib.plsvo_oracle_seed_update_batch.restype = C.c_int
    sbb, keep_sd = abi.make_seed_batch(sdd)
    cpu_sd = abi.SeedOut(ns)
    rate, th = best_threads(lambda t: olib.plsvo_oracle_seed_update_batch(C.byref(sbb), C.byref(cpu_sd.struct), t), ns)
    up = cpu_sd.status == abi.SEED_UPDATED
    exact = bool(np.array_equal(dfo["gpu"].status, cpu_sd.status) and np.array_equal(dfo["gpu"].depth[up], cpu_sd.depth[up]))
    okm = up & np.isfinite(cpu_sd.sigma2)
    res["depth_filter_seed_update"] = {
        "workload": f"{ns} point seeds, 8 keyframes -> 8 current VGA frames, 3 pyramid levels, 20 % edgelets "
                    "(DepthFilter::updatePointSeeds body: visibility, findEpipolarMatchDirect, computeTau, updatePointSeed)",
        "kernel_ms": k_ms, "e2e_ms": e_ms, "seeds_per_s_kernel": ns / (k_ms * 1e-3), "seeds_per_s_e2e": ns / (e_ms * 1e-3),
        "status_and_depth_bit_exact_vs_oracle": exact,
        "max_rel_diff_mu": float(np.max(np.abs(dfo["gpu"].mu[okm] - cpu_sd.mu[okm]) / np.abs(cpu_sd.mu[okm]))) if okm.any() else None,
        "status_hist_not_visible_no_match_updated": np.bincount(cpu_sd.status, minlength=3).tolist(),
        "cpu_oracle_seeds_per_s": rate, "cpu_threads": th}
    # ---- line seeds -------------------------------------------------------------------------------------------
    ls = pin_obj(synth.make_line_seed_batch(n=ns // 2, n_ref=8, n_cur=8, n_pyr_levels=3, seed=6, device=dev))
    lo = {}

    def run_ls():
        lo["gpu"] = dfilt.updateLineSeeds(ls)

    k_ms, e_ms = timed(ctx, run_ls, args.reps)
    olib.plsvo_oracle_line_seed_update_batch.restype = C.c_int
    lbb, keep_ls = abi.make_line_seed_batch(ls)
    cpu_ls = abi.LineSeedOut(ls.n)
    rate, th = best_threads(lambda t: olib.plsvo_oracle_line_seed_update_batch(C.byref(lbb), C.byref(cpu_ls.line_struct), t), ls.n)
    upl = cpu_ls.status == abi.SEED_UPDATED
    exact = bool(np.array_equal(lo["gpu"].status, cpu_ls.status) and np.array_equal(lo["gpu"].depth[upl], cpu_ls.depth[upl])
                 and np.array_equal(lo["gpu"].depth_e[upl], cpu_ls.depth_e[upl]))
    res["depth_filter_line_seed_update"] = {
        "workload": f"{ls.n} line seeds (two end-point searches each), 8 keyframes -> 8 current VGA frames (DepthFilter::updateLineSeeds body)",
        "kernel_ms": k_ms, "e2e_ms": e_ms, "seeds_per_s_kernel": ls.n / (k_ms * 1e-3), "seeds_per_s_e2e": ls.n / (e_ms * 1e-3),
        "status_and_depths_bit_exact_vs_oracle": exact,
        "status_hist_not_visible_no_match_updated": np.bincount(cpu_ls.status, minlength=3).tolist(),
        "cpu_oracle_seeds_per_s": rate, "cpu_threads": th}
    res["host_memory"] = "pinned (page-locked) inputs for every e2e_ms"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--features", type=int, default=400000)
    args = ap.parse_args()
    print(json.dumps(measure(args.reps, args.images, args.features), indent=1))


if __name__ == "__main__":
    main()
