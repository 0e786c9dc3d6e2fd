P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 2097152 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refine2 -s 2 -c 1 -o gpurun_out/prof_refine_r02 $P1 > gpurun_out/ncu_refine.log 2>&1; echo ncu_exit=$?
