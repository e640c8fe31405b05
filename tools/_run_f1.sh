mkdir -p gpurun_out; rm -f gpurun_out/ab_f1.txt
PAR_SEEDS=2 PAR_B=512 PAR_TIMEOUT=150 bash tools/gpu_step.sh f1 par | tail -2
TUNE_CONFIGS='[{"PLSVO_VARIANT":"128,4"},{"PLSVO_VARIANT":"256,2"},{"PLSVO_VARIANT":"96,5"}]' timeout 150 python tools/tune.py >> gpurun_out/ab_f1.txt 2>&1
cat gpurun_out/ab_f1.txt
timeout 300 python -m pytest tests/test_gpu_align.py tests/test_gpu_track.py -q --timeout 150 -k "not campaign" 2>&1 | tail -3
