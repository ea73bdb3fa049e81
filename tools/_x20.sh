set -u
timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128.*cluster=2" 
TIMEGAN_B200_CLUSTER_NO=2 timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128 B=256 cluster=2"
timeout 900 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "cluster" 2>&1 | tail -3
