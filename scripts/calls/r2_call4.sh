# round 2, call 4: consolidated state (FP16 tensor engine, queue/drain hit path, host staging, big k, new bench)
set -x
timeout 120 python scripts/dbg_bigk.py 2>&1 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/fuzz_parity.py 200 31 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -3 gpurun_out/fuzz.log
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo bench_exit=$?; tail -c 9000 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo ref_exit=$?; tail -c 2500 gpurun_out/bench_ref.log
