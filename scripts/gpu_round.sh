set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "c4_shape" --durations=3 2>&1 | tail -15
