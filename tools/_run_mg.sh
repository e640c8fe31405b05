N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; cut -c1-300 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
if [ "$N" = "8" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config c4 --batch 1024 --steps 5 --warmup 3 > gpurun_out/bench_c4_${N}gpu.json 2> gpurun_out/bench_c4_${N}gpu.err; cut -c1-300 gpurun_out/bench_c4_${N}gpu.json; tail -2 gpurun_out/bench_c4_${N}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --sweep --batch 1024 --steps 6 --sweep-out gpurun_out/sweep_${N}gpu.json > gpurun_out/sweep_${N}gpu.out 2> gpurun_out/sweep_${N}gpu.err; tail -c 300 gpurun_out/sweep_${N}gpu.err
fi
