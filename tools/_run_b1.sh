mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_track.py tests/test_gpu_align.py tests/test_gpu_abi_errors.py tests/test_gpu_shim.py -q --timeout 120 > gpurun_out/pytest_b1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b1.log; tail -25 gpurun_out/pytest_b1.log

import json; d=json.load(open('gpurun_out/bench_b1.json')); print('value',d['value'],'e2e',d['e2e'],'frac',d['roofline']['frac']); print('poseopt',json.dumps(d['poseopt'])[:900])"; tail -3 gpurun_out/bench_b1.err
