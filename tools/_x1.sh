set -u
O=gpurun_out
python -m pytest tests/test_steps_gpu.py -q -m gpu -k "graph or joint" 2>&1 | tail -3 > $O/x1_pytest.log; tail -2 $O/x1_pytest.log
TIMEGAN_B200_OVERLAP_STEPS=0 python bench.py --no-cpu-baseline > $O/x1_bench_off.json 2> $O/x1_bench_off.err; echo rc=$?
TIMEGAN_B200_OVERLAP_STEPS=1 python bench.py --no-cpu-baseline > $O/x1_bench_on.json 2> $O/x1_bench_on.err; echo rc=$?
python - <<'P'
import json
for f in ("off","on"):
    try:
        d=json.loads(open(f"gpurun_out/x1_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("also",{}).get("c3",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
P
