mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_track.py tests/test_gpu_align.py -q --timeout 200 -k "not campaign" > gpurun_out/pytest_e1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e1.log; tail -5 gpurun_out/pytest_e1.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e1.json 2> gpurun_out/bench_e1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_e1.json') if l.startswith('{')][0]); print('value',d['value'],'e2e',d['e2e'])"; tail -3 gpurun_out/bench_e1.err
