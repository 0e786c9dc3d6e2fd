SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_rstats.so PYTHONPATH=$PWD timeout 600 python scripts/retry_stats.py 2>&1 | tail -5
