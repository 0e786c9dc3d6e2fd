set -x
for g in fused copy; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 --only c3 --no-e2e --no-est --no-peaks --gather $g > gpurun_out/bench_2gpu_$g.log 2> gpurun_out/bench_2gpu_$g.err; echo "$g exit=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_2gpu_$g.log") if l.startswith("{")][-1])
    print("$g: N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "gather_ok", d["fused_gather_verified"], d.get("gather"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_2gpu_$g.err").read()[-2000:])
PY
done
timeout 600 python bench.py --steps 8 --warmup 3 --only c3 --no-e2e --no-est --no-peaks --no-cpu-baseline > gpurun_out/bench_1gpu_ref.log 2>&1; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_1gpu_ref.log") if l.startswith("{")][-1])
print("1 GPU same box: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"])
PY
