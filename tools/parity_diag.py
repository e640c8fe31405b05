"""Distribution of GPU-vs-oracle pose differences on a synthetic batch (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import plsvo_b200
from plsvo_b200 import abi, synth
import oracle_lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n_segs = int(sys.argv[3]) if len(sys.argv) > 3 else 80
data = synth.make_align_batch(batch=B, n_pts=n_pts, n_segs=n_segs, device="cuda", seed=4242)
gpu = plsvo_b200.SparseImgAlign(4, 2, 30).run(data)
ref = oracle_lib.align(abi, data, n_threads=32)
ang, rel = synth.pose_error(gpu.T_cur_w, ref.T_cur_w)
same = (gpu.iters == ref.iters).all(axis=1)
q = [50, 90, 99, 100]
print(os.environ.get("PLSVO_LIB", "default"), f"B={B} pts={n_pts} segs={n_segs}")
print("  same-iteration pairs: %.1f%%" % (100 * same.mean()))
print("  rot  pct", q, np.percentile(ang, q))
print("  relt pct", q, np.percentile(rel, q))
if same.any():
    print("  (same iters) rot max %.2e relt max %.2e" % (ang[same].max(), rel[same].max()))
print("  within tol: %d / %d" % (((ang <= 1e-5) & (rel <= 1e-4)).sum(), B))
