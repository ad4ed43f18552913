python tools/sweep_bench.py 200000 500000 3,4,8,16,31
python tools/sweep_bench.py 200000 62500 1,4,8,31
python tools/sweep_bench.py 200000 125000 4,31
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rhs_limbs.py tests/test_gpu_scale_parity.py -m gpu -x -q 2>&1 | tail -3)
