set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "hamming or gbnn or gb_forest or rfnn or forest" 2>&1 | tail -25
