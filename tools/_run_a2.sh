mkdir -p gpurun_out
for lib in "" tree fp32 tree_fp32; do
  if [ -n "$lib" ]; then export PLSVO_LIB=$PWD/pl-svo_b200/csrc/libplsvo_b200_$lib.so; else unset PLSVO_LIB; fi
  echo "== lib ${lib:-default}" >> gpurun_out/ab_a2.txt
  TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"},{"PLSVO_VARIANT":"96,5"}]' timeout 300 python tools/tune.py >> gpurun_out/ab_a2.txt 2>&1
done
unset PLSVO_LIB
cat gpurun_out/ab_a2.txt
TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"}]' timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_img_align -s 2 -c 1 -o gpurun_out/prof_a2 python tools/tune.py > gpurun_out/ncu_a2.log 2>&1
tail -3 gpurun_out/ncu_a2.log
