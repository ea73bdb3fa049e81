set -u
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 > $O/r02f_bench_dp8.json 2> $O/r02f_bench_dp8.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02f_bench_dp8.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("dp_check"), d.get("also",{}).get("c3",{}).get("value"), d.get("also",{}).get("c3",{}).get("ms_per_step"))
P
