set -u
O=gpurun_out
timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128" 
TIMEGAN_B200_CLUSTER_DIO=0 timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128 B=256 cluster=2"
timeout 900 python -m pytest tests/test_gru_gpu.py tests/test_bench_shape_gpu.py -q -m gpu -k "cluster or 128 or c3" 2>&1 | tail -3
