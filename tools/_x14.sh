set -u
O=gpurun_out
for cfg in "1 0" "1 1" "2 1" "2 0"; do
set -- $cfg
TIMEGAN_B200_STREAM_PRIO=$1 TIMEGAN_B200_SLACK_STREAM=$2 python bench.py --no-cpu-baseline --no-also-c3 > $O/x14_bench_$1_$2.json 2> $O/x14_bench_$1_$2.err; echo rc=$?
done
python - <<'P'
import json
for f in ("1_0","1_1","2_1","2_0"):
    try:
        d=json.loads(open(f"gpurun_out/x14_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
P
TIMEGAN_B200_STREAM_PRIO=2 TIMEGAN_B200_SLACK_STREAM=1 python -m pytest tests/test_steps_gpu.py -q -m gpu -k "graph" 2>&1 | tail -2
