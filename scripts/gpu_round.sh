set -x
B="python bench.py --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_ll.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v11.csv $B > gpurun_out/ncu_ll.log 2>&1; tail -1 gpurun_out/ncu_ll.log | cut -c1-200
