set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "integration_stub" 2>&1 | tail -15
