python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 80 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
