"""One-process tuning sweep over the alignment kernel's compiled variants (env knobs read at launch time)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import plsvo_b200
from plsvo_b200 import abi, synth

B = int(os.environ.get("TUNE_B", 1024)); n_pts = int(os.environ.get("TUNE_PTS", 300)); n_segs = int(os.environ.get("TUNE_SEGS", 80))
dev = torch.device("cuda", 0)
data = synth.make_align_batch(batch=B, n_pts=n_pts, n_segs=n_segs, device=dev, seed=3000)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = plsvo_b200.Context(0, stream.cuda_stream)
al = plsvo_b200.SparseImgAlign(4, 2, 30, ctx=ctx)
al.upload(data)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
KEYS = ("PLSVO_VARIANT", "PLSVO_CTAS_PER_SM", "PLSVO_IMG_SMEM")
configs = json.loads(os.environ.get("TUNE_CONFIGS", "[]")) or [
    {"PLSVO_VARIANT": v} for v in ("128,4", "128,5", "96,5", "96,7", "64,8", "256,2")]
base = None
for cfg in configs:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in cfg.items()})
    try:
        for _ in range(2):
            al.launch()
        ctx.sync()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); al.launch(); e.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(s.elapsed_time(e))
        out = al.download()
        if base is None:
            base = out
        ang, rel = synth.pose_error(out.T_cur_w, base.T_cur_w)
        print(json.dumps({"cfg": cfg, "ms": round(float(np.median(ts)), 4), "pairs_per_s": round(B / (np.median(ts) * 1e-3)),
                          "vs_first_max_rot": float(ang.max()), "same_iters": float((out.iters == base.iters).all(axis=1).mean()),
                          "flags": int((out.status >> 2).astype(bool).sum())}), flush=True)
    except Exception as ex:
        print(json.dumps({"cfg": cfg, "error": str(ex)}), flush=True)
