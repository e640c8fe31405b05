mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_track.py -q -s -k "sequence" --timeout 250 > gpurun_out/pytest_b3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b3.log; grep -E "drift|passed|failed|Error" gpurun_out/pytest_b3.log | head
timeout 200 python tools/b1_latency.py 200 > gpurun_out/b1_latency.txt 2> gpurun_out/b1_latency.err; cat gpurun_out/b1_latency.txt; tail -2 gpurun_out/b1_latency.err
timeout 240 ncu --set full --clock-control none --import-source on -k regex:sparse_img_align -s 3 -c 1 -o gpurun_out/prof_b3 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_b3.log 2>&1; tail -2 gpurun_out/ncu_b3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_b3.log 2>&1; grep -c "sparse_img\|pose_opt" gpurun_out/launches_b3.csv
timeout 300 python tools/parity_campaign.py 8 1024 > gpurun_out/parity_8192.txt 2>&1; tail -2 gpurun_out/parity_8192.txt
