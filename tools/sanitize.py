"""Small runs of every alignment / tracking path for compute-sanitizer (memcheck, racecheck, synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize.py"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import plsvo_b200
from plsvo_b200 import abi, synth
import oracle_lib

def check(tag, gpu, ref):
    ang, rel = synth.pose_error(gpu.T_cur_w, ref.T_cur_w)
    print(tag, "iters equal", bool((gpu.iters == ref.iters).all()), "max rot %.2e" % ang.max(), "status", sorted(set(gpu.status.tolist())), flush=True)

d = synth.make_align_batch(batch=3, n_pts=150, n_segs=30, device="cuda", seed=11)
check("vga 150+30      ", plsvo_b200.SparseImgAlign(4, 2, 30).run(d), oracle_lib.align(abi, d))
d = synth.make_align_batch(cam=synth.HD720, batch=2, n_pts=200, n_segs=12, max_level=2, min_level=0, device="cuda", seed=3900, motion_t=0.004, motion_r=0.0012, margin=24)
check("720p levels 2->0", plsvo_b200.SparseImgAlign(2, 0, 30).run(d), oracle_lib.align(abi, d, abi.align_params(2, 0, 30)))
d = synth.make_align_batch(batch=300, n_pts=40, n_segs=8, device="cuda", seed=12)
lean = copy.copy(d); lean.ref_pyr = {2: d.ref_pyr[2]}; lean.cur_pyr = {2: d.cur_pyr[2]}; lean.pt_f = lean.seg_sf = lean.seg_ef = None
check("gated, derived  ", plsvo_b200.SparseImgAlign(4, 2, 30).run(lean), oracle_lib.align(abi, d, n_threads=8))
os.environ["PLSVO_VARIANT"] = "256,2"
d = synth.make_align_batch(batch=2, n_pts=300, n_segs=80, device="cuda", seed=13)
check("variant 256,2   ", plsvo_b200.SparseImgAlign(4, 2, 30).run(d), oracle_lib.align(abi, d))
os.environ.pop("PLSVO_VARIANT")
al, po = synth.make_track_batch(batch=3, n_pts=120, n_segs=24, seed=14, device="cuda")
ao, pout = plsvo_b200.api.track(al, po)
print("track ok", ao.iters.sum(), pout.num_obs_pt.tolist(), flush=True)
# next-row kernels whose outputs grew in round 2 (Matcher::A_cur_ref_, end-point px_cur of the line seeds)
md = synth.make_match_batch(n=300, seed=15, device="cuda")
mo = plsvo_b200.api.Matcher(10).findMatchDirect(md)
mr = oracle_lib.match_direct(abi, md, 4)
print("match_direct A equal", bool(np.array_equal(mo.A_cur_ref, mr.A_cur_ref, equal_nan=True)), "success equal", bool(np.array_equal(mo.success, mr.success)), flush=True)
ls = synth.make_line_seed_batch(n=200, seed=16, device="cuda")
lo = plsvo_b200.api.DepthFilter().updateLineSeeds(ls)
lr = oracle_lib.line_seed_update(abi, ls, 4)
print("line seeds status equal", bool(np.array_equal(lo.status, lr.status)), "px_cur_e equal", bool(np.array_equal(lo.px_cur_e, lr.px_cur_e, equal_nan=True)), flush=True)
