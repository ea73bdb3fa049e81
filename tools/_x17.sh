set -u
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > $O/r02f_bench_dp8.json 2> $O/r02f_bench_dp8.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02f_bench_dp8.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("dp_check"), d.get("also",{}).get("c3",{}))
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/trace_overlap.py --out $O/r02f_timeline_c2_8gpu.csv > $O/r02f_timeline_c2_8gpu.log 2>&1; echo rc=$?; grep timeline $O/r02f_timeline_c2_8gpu.log | head -30
