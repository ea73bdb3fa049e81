set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gru_gpu.py -q -m gpu -k "bf16" 2>&1 | tail -25 > $O/x4_pytest_a.log; tail -12 $O/x4_pytest_a.log
python bench.py --no-cpu-baseline --no-also-c3 --hidden 128 --proj bf16 --steps 8 --warmup 3 > $O/x4_bench_c3_bf16.json 2> $O/x4_bench_c3_bf16.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/x4_bench_c3_bf16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["config"]["workload"])
print({k:(v["ms_per_step"],v["GBps"]) for k,v in d["families"].items()})
P
