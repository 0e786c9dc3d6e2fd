set -x
for s in 2 3 4; do timeout 900 python scripts/fuzz_parity.py 150 $s 2>&1 | tail -2; done
