set -x
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo bench8_exit=$?; grep -v "^\*\|OMP_NUM" gpurun_out/bench_8gpu.err | tail -5 | cut -c1-500; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_8gpu.log") if l.startswith("{")][-1])
    print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], "gather_ok", d["fused_gather_verified"])
    print("c5", d["c5"]["value"], d["c5"]["ms_per_step"], "c4", d["c4"]["value"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_8gpu.log").read()[-3000:])
PY
