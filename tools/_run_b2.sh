mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_b2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b2.log; tail -12 gpurun_out/pytest_b2.log
timeout 500 python bench.py --sweep --batch 1024 --steps 6 --sweep-out gpurun_out/sweep_1gpu.json > gpurun_out/sweep_b2.out 2> gpurun_out/sweep_b2.err; tail -c 600 gpurun_out/sweep_b2.err; grep -c n_pts gpurun_out/sweep_b2.err
