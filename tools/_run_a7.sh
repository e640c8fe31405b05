mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_a7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_a7.log; tail -15 gpurun_out/pytest_a7.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_a7.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_a7.log; tail -4 gpurun_out/smoke_a7.log
for lib in "" prefetch; do
  if [ -n "$lib" ]; then export PLSVO_LIB=$PWD/pl-svo_b200/csrc/libplsvo_b200_$lib.so; else unset PLSVO_LIB; fi
  echo "== lib ${lib:-default}" >> gpurun_out/ab_a7.txt
  TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"}]' timeout 120 python tools/tune.py >> gpurun_out/ab_a7.txt 2>&1
done
unset PLSVO_LIB
cat gpurun_out/ab_a7.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a7.json 2> gpurun_out/bench_a7.err; cut -c1-2500 gpurun_out/bench_a7.json; tail -3 gpurun_out/bench_a7.err
