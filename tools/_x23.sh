TIMEGAN_B200_CLUSTER_DIO=1 timeout 300 python tools/probe_cluster.py 2>&1 | grep "H=128.*cluster=2"
TIMEGAN_B200_CLUSTER_DIO=1 timeout 300 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "cluster" 2>&1 | tail -2
