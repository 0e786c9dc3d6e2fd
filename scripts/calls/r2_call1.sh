# round 2, call 1: fp16 tensor engine - correctness + speed
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/fuzz_parity.py 150 21 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -3 gpurun_out/fuzz.log
for ss in 4 2 8; do
timeout 300 python bench.py --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-e2e --tc-seed-stride $ss > gpurun_out/bench_f16_s$ss.log 2>&1; echo exit=$?; tail -c 1500 gpurun_out/bench_f16_s$ss.log | grep -o '"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
done
timeout 300 python bench.py --steps 3 --warmup 2 --n-queries 4194304 --dim 64 --no-cpu-baseline --no-e2e > gpurun_out/bench_f16_d64.log 2>&1; echo exit=$?; tail -c 1500 gpurun_out/bench_f16_d64.log | grep -o '"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f16_full.log 2>&1; echo exit=$?; tail -c 3000 gpurun_out/bench_f16_full.log
# ---- variant q: queue/drain hit path + joint threshold ----
export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_q.so
if [ -f "$SKNNR_B200_LIB" ]; then
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke q ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo pytest_q_exit=$?; tail -5 gpurun_out/pytest_q.log
timeout 600 python scripts/fuzz_parity.py 150 22 > gpurun_out/fuzz_q.log 2>&1; echo fuzz_q_exit=$?; tail -3 gpurun_out/fuzz_q.log
for ss in 4 2 8; do
timeout 300 python bench.py --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-e2e --tc-seed-stride $ss > gpurun_out/bench_q_s$ss.log 2>&1; echo exit=$?; tail -c 1500 gpurun_out/bench_q_s$ss.log | grep -o '"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
done
timeout 300 python bench.py --steps 3 --warmup 2 --n-queries 4194304 --dim 64 --no-cpu-baseline --no-e2e > gpurun_out/bench_q_d64.log 2>&1; echo exit=$?; tail -c 1500 gpurun_out/bench_q_d64.log | grep -o '"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
timeout 300 python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e --tc-debug 8 > gpurun_out/bench_q_cnt.log 2>&1; grep "tc counters" gpurun_out/bench_q_cnt.log | tail -1
fi
