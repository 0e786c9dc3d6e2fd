python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "host_pipeline or pageable or finite or full_size" 2>&1 | tail -2
