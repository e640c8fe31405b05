mkdir -p gpurun_out
timeout 200 python tools/tune_e2e.py > gpurun_out/tune_e2e_c2.txt 2>&1; cat gpurun_out/tune_e2e_c2.txt
