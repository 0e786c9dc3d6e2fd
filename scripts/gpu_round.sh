set -x
B="python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tc10 $B > gpurun_out/ncu_full.log 2>&1; echo ncu_full_exit=$?
timeout 300 $B > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:refine -s 2 -c 1 -o gpurun_out/prof_refine3 $B > gpurun_out/ncu_full2.log 2>&1; echo ncu_full2_exit=$?
B4="python bench.py --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-e2e"
timeout 300 $B4 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_tc10.csv $B4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
