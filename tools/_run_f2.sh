mkdir -p gpurun_out; rm -f gpurun_out/ab_f2.txt
TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"},{"PLSVO_VARIANT":"160,3"},{"PLSVO_VARIANT":"192,2"},{"PLSVO_VARIANT":"256,2"}]' timeout 200 python tools/tune.py >> gpurun_out/ab_f2.txt 2>&1
cat gpurun_out/ab_f2.txt
