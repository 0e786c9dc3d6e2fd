set -x
cat > /tmp/hb.py <<'PY'
import sys; sys.path.insert(0, '.')
from scripts.hamming_bench import main
main(n_q=200000, weighted=("w" in sys.argv))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_search_kernel -s 1 -c 1 -o gpurun_out/prof_ham_u -f python /tmp/hb.py > gpurun_out/ncu_ham_u.log 2>&1; tail -2 gpurun_out/ncu_ham_u.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_search_kernel -s 1 -c 1 -o gpurun_out/prof_ham_w -f python /tmp/hb.py w > gpurun_out/ncu_ham_w.log 2>&1; tail -2 gpurun_out/ncu_ham_w.log
