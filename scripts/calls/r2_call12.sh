set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "certifies_nearly" 2>&1 | tail -4
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $B > gpurun_out/bench_d.log 2>&1; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_d.log") if l.startswith("{")][-1])
print("value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
timeout 900 python scripts/fuzz_parity.py 150 51 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -2 gpurun_out/fuzz.log
P="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $P > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv $P > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $P1 > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tc_r02 $P1 > gpurun_out/ncu_tc.log 2>&1; echo ncu_exit=$?
