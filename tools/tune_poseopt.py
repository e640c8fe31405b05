import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import plsvo_b200
from plsvo_b200 import abi, synth
B = int(os.environ.get("TUNE_B", 4096))
dev = torch.device("cuda", 0)
pdata = synth.make_poseopt_batch(batch=B, n_pts=300, n_segs=80, seed=5000)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = plsvo_b200.Context(0, stream.cuda_stream)
pbatch, keep = abi.make_poseopt_batch(pdata)
pparams = abi.poseopt_params(2.0, 10, -1)
ctx.check(ctx.lib.plsvo_poseopt_upload(ctx.handle, C.byref(pbatch)), "up")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): ctx.check(ctx.lib.plsvo_poseopt_launch(ctx.handle, C.byref(pparams)), "l")
ts = []
for _ in range(7):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream); ctx.check(ctx.lib.plsvo_poseopt_launch(ctx.handle, C.byref(pparams)), "l"); e.record(stream)
    torch.cuda.synchronize(dev); ts.append(s.elapsed_time(e))
print(json.dumps({"lib": os.environ.get("PLSVO_LIB", "default"), "ms": round(float(np.median(ts)), 4), "frames_per_s": round(B / (np.median(ts) * 1e-3))}))
