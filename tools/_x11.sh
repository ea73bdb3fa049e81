set -u
O=gpurun_out
timeout 300 python tools/probe_cluster.py > $O/x11_probe_cluster.log 2>&1; echo rc=$?; cat $O/x11_probe_cluster.log | tail -12
timeout 600 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "cluster or tangent" 2>&1 | tail -3
