mkdir -p gpurun_out; rm -f gpurun_out/ab_c1.txt
PAR_SEEDS=2 PAR_B=512 PAR_TIMEOUT=150 bash tools/gpu_step.sh c1 par | tail -3
for lib in "" evalshared; do
  if [ -n "$lib" ]; then export PLSVO_LIB=$PWD/pl-svo_b200/csrc/libplsvo_b200_$lib.so; else unset PLSVO_LIB; fi
  echo "== lib ${lib:-default}" >> gpurun_out/ab_c1.txt
  TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"}]' timeout 120 python tools/tune.py >> gpurun_out/ab_c1.txt 2>&1
done
unset PLSVO_LIB
cat gpurun_out/ab_c1.txt
