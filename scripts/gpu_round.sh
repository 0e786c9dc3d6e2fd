# One round on the GPU box: `gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'`
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
