set -u
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r02f_pytest_gpu.log; tail -2 $O/r02f_pytest_gpu.log
python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$?"
python bench.py --no-cpu-baseline --no-also-c3 --hidden 128 --proj bf16 --steps 10 --warmup 3 > $O/r02f_bench_c3_bf16.json 2> $O/r02f_bench_c3.err; echo rc=$?
python tools/sweep_phases.py --hidden 24 64 128 256 --steps 6 > $O/r02f_sweep_phases_c4_1gpu.jsonl 2>$O/r02f_sweep.err; echo "c4 rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["also"]["c3"]["ms_per_step"], d["cpu_baseline"]["value"])
d=json.loads(open("gpurun_out/r02f_bench_c3_bf16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d["families"].items()})
for l in open("gpurun_out/r02f_sweep_phases_c4_1gpu.jsonl"):
    j=json.loads(l); print(j["hidden"], j["ae_ms"], j["sup_ms"], j["joint_ms"])
P
