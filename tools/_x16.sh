set -u
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $O/r02f_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also-c3 --no-graph > $O/r02f_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/sweep_phases.py --hidden 24 64 128 256 --steps 6 > $O/r02f_sweep_phases_c4_1gpu.jsonl 2>$O/r02f_sweep.err; echo "c4 rc=$?"; cut -c1-260 $O/r02f_sweep_phases_c4_1gpu.jsonl
python bench.py --no-cpu-baseline --no-also-c3 --hidden 128 --proj bf16 --steps 10 --warmup 3 > $O/r02f_bench_c3_bf16.json 2> $O/r02f_bench_c3.err; echo rc=$?
python -m pytest tests/test_bench_shape_gpu.py tests/test_steps_gpu.py -q -m gpu -k "bf16 or tf32" 2>&1 | tail -2
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02f_bench_c3_bf16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d["families"].items()})
P
