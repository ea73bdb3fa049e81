set -u
O=gpurun_out
timeout 300 python tools/probe_jvp256.py > $O/r02_probe_jvp256.log 2>&1; echo rc=$?; cat $O/r02_probe_jvp256.log | tail -8
timeout 600 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "tangent or capacity" 2>&1 | tail -4
