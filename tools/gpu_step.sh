#!/bin/bash
# One GPU-box visit while developing: sanitizer on a tiny run, parity over seeded pairs, variant sweep, GPU tests.
# Usage (under gpurun, from the repo root): bash tools/gpu_step.sh <tag> [steps...]   steps: san par tune test bench
tag=${1:-dev}; shift
steps=${@:-san par tune test}
mkdir -p gpurun_out
for s in $steps; do
  case $s in
    san) timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path[:0]=['.','oracle']
import plsvo_b200, numpy as np
from plsvo_b200 import abi, synth
import oracle_lib
d = synth.make_align_batch(batch=3, n_pts=150, n_segs=30, device='cuda', seed=11)
g = plsvo_b200.SparseImgAlign(4,2,30).run(d)
r = oracle_lib.align(abi, d)
print('iters gpu', g.iters.tolist(), 'ref', r.iters.tolist(), 'status', g.status.tolist())
print('pose err', synth.pose_error(g.T_cur_w, r.T_cur_w))
" > gpurun_out/san_$tag.log 2>&1; echo "san rc=$?" >> gpurun_out/san_$tag.log; tail -12 gpurun_out/san_$tag.log;;
    par) timeout ${PAR_TIMEOUT:-240} python tools/parity_campaign.py ${PAR_SEEDS:-2} ${PAR_B:-512} > gpurun_out/parity_$tag.txt 2>&1; echo "par rc=$?" >> gpurun_out/parity_$tag.txt; tail -25 gpurun_out/parity_$tag.txt;;
    tune) timeout 200 python tools/tune.py > gpurun_out/tune_$tag.txt 2>&1; echo "tune rc=$?" >> gpurun_out/tune_$tag.txt; tail -12 gpurun_out/tune_$tag.txt;;
    test) timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -15 gpurun_out/pytest_$tag.log;;
    testall) timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -25 gpurun_out/pytest_$tag.log;;
    bench) timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-3000 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err;;
  esac
done
