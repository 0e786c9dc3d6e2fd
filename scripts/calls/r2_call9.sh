set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "pageable or finite or beyond_shared or non_integer or nodata_is" 2>&1 | tail -8
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks"
timeout 300 $B > gpurun_out/bench_est1.log 2>&1; grep -o '"e2e_estimator": {[^}]*}' gpurun_out/bench_est1.log | cut -c1-600
timeout 300 python scripts/est_profile.py 4194304 2>&1 | head -3
timeout 300 python -c "
import torch, sys, time, numpy as np
sys.argv=['x','4194304']
torch.zeros(1).cuda()
exec(open('scripts/est_profile.py').read())
" 2>&1 | head -3
