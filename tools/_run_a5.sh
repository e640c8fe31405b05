mkdir -p gpurun_out
PAR_SEEDS=2 PAR_B=512 bash tools/gpu_step.sh a5 par
for lib in "" tree; do
  if [ -n "$lib" ]; then export PLSVO_LIB=$PWD/pl-svo_b200/csrc/libplsvo_b200_$lib.so; else unset PLSVO_LIB; fi
  echo "== lib ${lib:-default}" >> gpurun_out/ab_a5.txt
  TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"}]' timeout 120 python tools/tune.py >> gpurun_out/ab_a5.txt 2>&1
done
unset PLSVO_LIB
cat gpurun_out/ab_a5.txt
TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"}]' timeout 240 ncu --set full --clock-control none --import-source on -k regex:sparse_img_align -s 2 -c 1 -o gpurun_out/prof_a5 python tools/tune.py > gpurun_out/ncu_a5.log 2>&1
tail -2 gpurun_out/ncu_a5.log
