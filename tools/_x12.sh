set -u
O=gpurun_out
for m in tf32 bf16; do
python bench.py --no-cpu-baseline --no-also-c3 --hidden 128 --proj $m --steps 10 --warmup 3 > $O/x12_bench_c3_$m.json 2> $O/x12_bench_c3_$m.err; echo rc=$?
done
python - <<'P'
import json
for m in ("tf32","bf16"):
    d=json.loads(open(f"gpurun_out/x12_bench_c3_{m}.json").read().strip().splitlines()[-1])
    print(m, d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in d["families"].items()})
P
