set -u
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > $O/x8_pytest_all.log; tail -3 $O/x8_pytest_all.log
