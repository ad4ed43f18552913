(timeout 600 python -m pytest tests/test_gpu_dense_grm.py -m gpu -x -q 2>&1 | tail -5)
timeout 300 python tools/dense_bench.py 20000 100000 7 full 2>&1 | tail -6
timeout 300 python tools/dense_bench.py 40000 50000 7 full 2>&1 | tail -6
