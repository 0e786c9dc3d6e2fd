timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -1 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 200 93 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
B10="python bench.py --only c3 --steps 3 --warmup 2 --no-cpu-baseline --no-peaks --no-est --no-e2e"
timeout 600 $B10 > gpurun_out/b10.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b10.log") if l.startswith("{")][-1])
print("10M: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"])
PY
P4="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_r02c.csv $P4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
grep -E "refine2" gpurun_out/launches_r02c.csv | awk -F'","' '{print $(NF-2), $NF}' | head -8
