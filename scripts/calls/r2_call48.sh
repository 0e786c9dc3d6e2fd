timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --only c3 --no-e2e --no-est --no-peaks > gpurun_out/bench_2gpu_ab.log 2> gpurun_out/bench_2gpu_ab.err; echo exit=$?; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu_ab.err | tail -3 | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_2gpu_ab.log") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "gather", d["gather"], d["fused_gather_verified"], "alt", d["gather_alt"])
PY
