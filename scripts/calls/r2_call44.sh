P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-peaks --no-e2e --no-est"
P4="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $P1 > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tc_r02f -f $P1 > gpurun_out/ncu_tc.log 2>&1; echo ncu_exit=$?
timeout 300 $P4 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02f.csv $P4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
ls -la gpurun_out/prof_tc_r02f.ncu-rep gpurun_out/launches_r02f.csv
