set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -6
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo bench2_exit=$?; tail -5 gpurun_out/bench_2gpu.err | cut -c1-700; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_2gpu.log") if l.startswith("{")][-1])
    print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], "gather_ok", d["fused_gather_verified"])
    print("c5", d["c5"]["value"], d["c5"]["ms_per_step"], "c4", d["c4"]["value"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_2gpu.log").read()[-3000:])
PY
SKNNR_B200_DEVICES=0,1 timeout 600 python scripts/est_profile.py 20000000 2>&1 | sed -n '1p'
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-sample 100000 2>&1 | tail -2 | cut -c1-400
