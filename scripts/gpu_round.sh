set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/pytest_gpu.log
for kc in 8 16; do
timeout 600 python bench.py --steps 2 --warmup 1 --n-queries 2097152 --no-cpu-baseline --kc $kc > gpurun_out/bench_kc$kc.log 2>&1; echo kc${kc}_exit=$?; tail -c 1800 gpurun_out/bench_kc$kc.log
done
