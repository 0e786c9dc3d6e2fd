timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > gpurun_out/pytest_2p.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_2p.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo bench2_exit=$?; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu.err | tail -3 | cut -c1-400
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_2gpu.log") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", (d.get("e2e_estimator") or {}).get("value"), "verified", d.get("fused_gather_verified"), "gather", d.get("gather"))
print("c5", d["c5"]["value"], d["c5"]["ms_per_step"], "c4", d["c4"]["value"])
PY
