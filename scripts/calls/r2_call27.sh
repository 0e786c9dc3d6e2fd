timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo pytest_exit=$?; tail -1 gpurun_out/pytest_k.log
B10="python bench.py --only c3 --steps 3 --warmup 2 --no-cpu-baseline --no-peaks --no-est --no-e2e"
timeout 600 $B10 > gpurun_out/b10.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b10.log") if l.startswith("{")][-1])
print("10M: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"])
PY
P4="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02c.csv $P4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
grep -E "project_kernel|refine2|search_simt" gpurun_out/launches_r02c.csv | awk -F'","' '{print $5, $NF}' | head -12
