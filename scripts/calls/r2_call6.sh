set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -8
timeout 600 python scripts/est_profile.py 2>&1 | tail -14
SKNNR_B200_PINNED_POOL_MB=0 timeout 600 python scripts/est_profile.py 2>&1 | head -3
