set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py -q -m gpu 2>&1 | tail -4 > $O/x9_pytest_peer.log; tail -2 $O/x9_pytest_peer.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_dp_graph.py > $O/x9_dp_graph.log 2>&1; echo rc=$?; grep "rank" $O/x9_dp_graph.log | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > $O/x9_bench_dp2.json 2> $O/x9_bench_dp2.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/x9_bench_dp2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("dp_check"), d.get("also",{}).get("c3",{}).get("ms_per_step"))
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/trace_overlap.py --out $O/r02_timeline_c2_2gpu.csv > $O/r02_timeline_c2_2gpu.log 2>&1; echo rc=$?; grep timeline $O/r02_timeline_c2_2gpu.log | head -30
