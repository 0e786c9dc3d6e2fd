set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-sample 200000 2>&1 | tail -1 | cut -c1-300
