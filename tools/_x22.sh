set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "cluster" 2>&1 | tail -3
timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128"
