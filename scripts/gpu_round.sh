set -x
timeout 600 python scripts/hamming_bench.py 2>&1 | grep -o "hamming.*id-compares/s\|RFNN knei.*queries/s"
timeout 900 python -m pytest tests -m gpu -x -q -k "hamming or gbnn or gb_forest or rfnn or forest or c4_shape" 2>&1 | tail -3
