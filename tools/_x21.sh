set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "cluster or tangent" 2>&1 | tail -3
timeout 300 python tools/probe_cluster.py > $O/r02f_probe_cluster.log 2>&1; grep "cluster=2" $O/r02f_probe_cluster.log
timeout 300 python tools/probe_jvp256.py 2>&1 | grep "cluster_jvp256=1"
python bench.py --no-cpu-baseline --no-also-c3 --hidden 128 --proj bf16 --steps 10 --warmup 3 > $O/r02f_bench_c3_bf16.json 2> $O/r02f_bench_c3.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02f_bench_c3_bf16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d["families"].items()})
P
