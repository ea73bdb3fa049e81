set -u
O=gpurun_out
for p in 1 0; do
TIMEGAN_B200_STREAM_PRIO=$p python bench.py --no-cpu-baseline --no-also-c3 > $O/x13_bench_p$p.json 2> $O/x13_bench_p$p.err; echo rc=$?
done
python - <<'P'
import json
for f in (1,0):
    d=json.loads(open(f"gpurun_out/x13_bench_p{f}.json").read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
P
