set -u
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/trace_overlap.py --out $O/r02f_timeline_c2_8gpu.csv > $O/r02f_timeline_c2_8gpu.log 2>&1; echo rc=$?; grep timeline $O/r02f_timeline_c2_8gpu.log | head -30
