set -x
timeout 600 python scripts/hamming_bench.py weighted 2>&1 | grep -o "hamming.*id-compares/s"
timeout 900 python -m pytest tests -m gpu -x -q -k "hamming or gbnn or gb_forest" 2>&1 | tail -3
