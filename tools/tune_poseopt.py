"""Pose-optimiser kernel time (BASELINE config 3: 300 points + 80 lines, <= 10 GN iterations, B = 4096 frames)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import plsvo_b200
from plsvo_b200 import abi, synth
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = plsvo_b200.Context(0, stream.cuda_stream)
pdata = synth.make_poseopt_batch(batch=int(os.environ.get("TUNE_B", 4096)), n_pts=300, n_segs=80, seed=5000)
pbatch, keep = abi.make_poseopt_batch(pdata)
pout = abi.PoseOptOut(pdata.batch, pdata.n_pts, pdata.n_segs)
pp = abi.poseopt_params(2.0, 10, -1)
ctx.check(ctx.lib.plsvo_poseopt_upload(ctx.handle, C.byref(pbatch)), "upload")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ctx.check(ctx.lib.plsvo_poseopt_launch(ctx.handle, C.byref(pp)), "launch")
torch.cuda.synchronize(dev)
ts = []
for _ in range(7):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream); ctx.check(ctx.lib.plsvo_poseopt_launch(ctx.handle, C.byref(pp)), "launch"); e.record(stream)
    torch.cuda.synchronize(dev); ts.append(s.elapsed_time(e))
ctx.check(ctx.lib.plsvo_poseopt_download(ctx.handle, C.byref(pout.struct)), "download")
import oracle_lib
ref = oracle_lib.poseopt(abi, pdata, pp, n_threads=32)
ang, rel = synth.pose_error(pout.T_f_w, ref.T_f_w)
print(json.dumps({"lib": os.environ.get("PLSVO_LIB", "default"), "ms": round(float(np.median(ts)), 4), "frames_per_s": round(pdata.batch / (np.median(ts) * 1e-3)),
                  "max_rot": float(ang.max()), "outliers_equal": bool(np.array_equal(pout.pt_outlier, ref.pt_outlier))}))
