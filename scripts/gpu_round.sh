set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "raster" 2>&1 | tail -25
