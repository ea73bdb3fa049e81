set -u
O=gpurun_out
for w in 0 40 74 148; do
python bench.py --no-cpu-baseline --no-also-c3 --wgrad-ctas $w > $O/x2_bench_w$w.json 2> $O/x2_bench_w$w.err; echo rc=$?
done
python - <<'P'
import json
for f in (0,40,74,148):
    try:
        d=json.loads(open(f"gpurun_out/x2_bench_w{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
P
