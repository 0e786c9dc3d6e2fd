set -x
timeout 600 python scripts/hamming_bench.py weighted 2>&1 | tail -4
