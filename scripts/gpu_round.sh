set -x
B="python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refine2_kernel -s 2 -c 1 -o gpurun_out/prof_refine4 -f $B > gpurun_out/ncu_refine4.log 2>&1; tail -1 gpurun_out/ncu_refine4.log
