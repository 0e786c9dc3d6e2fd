set -x
timeout 600 python scripts/raster_bench.py 2>&1 | tail -4
