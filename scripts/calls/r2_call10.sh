set -x
for n in 10000000 4194304 4000000; do timeout 300 python scripts/est_profile.py $n 2>&1 | sed -n '1p;4,6p'; done
