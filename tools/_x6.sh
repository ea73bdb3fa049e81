set -u
O=gpurun_out
python -m pytest tests/test_steps_gpu.py tests/test_train_gpu.py -q -m gpu -k "graph or fast or default" 2>&1 | tail -3 > $O/x6_pytest.log; tail -2 $O/x6_pytest.log
python bench.py --no-cpu-baseline > $O/x6_bench.json 2> $O/x6_bench.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/x6_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("also",{}).get("c3",{}).get("ms_per_step"))
P
python tools/trace_overlap.py --out $O/r02_timeline_c2_1gpu.csv > $O/r02_timeline_c2_1gpu.log 2>&1; echo rc=$?; grep timeline $O/r02_timeline_c2_1gpu.log | head -5
