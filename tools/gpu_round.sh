#!/bin/bash
# One GPU-box visit at the end of a round: GPU tests, smoke, the bench line, the reference arm, the ncu launch list of the
# same command (this repo's kernels only) and one `ncu --set full` capture of the alignment kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sparse_img_align|pose_optimizer|weight_selftest|pyramid' -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sparse_img_align -s 3 -c 1 -o gpurun_out/prof_$tag python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/pytest_$tag.log; tail -2 gpurun_out/smoke_$tag.log; cut -c1-1500 gpurun_out/bench_$tag.json; cut -c1-600 gpurun_out/bench_ref_$tag.json; grep -c sparse_img gpurun_out/launches_$tag.csv
