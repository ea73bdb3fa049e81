set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_gru_gpu.py -q -m gpu -x -k "bf16" 2>&1 | tail -25 > $O/x3_pytest_a.log; tail -12 $O/x3_pytest_a.log
timeout 900 python -m pytest tests/test_steps_gpu.py tests/test_bench_shape_gpu.py -q -m gpu -k "bf16 or tf32" 2>&1 | tail -25 > $O/x3_pytest_b.log; tail -12 $O/x3_pytest_b.log
