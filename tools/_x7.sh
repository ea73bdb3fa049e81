set -u
O=gpurun_out
for kb in 227 160 128 96; do
TIMEGAN_B200_GEMM_SMEM_KB=$kb python bench.py --no-cpu-baseline --no-also-c3 > $O/x7_bench_$kb.json 2> $O/x7_bench_$kb.err; echo rc=$?
done
python - <<'P'
import json
for f in (227,160,128,96):
    try:
        d=json.loads(open(f"gpurun_out/x7_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], {k:v["ms_per_step"] for k,v in d["families"].items() if k in ("proj","wgrad")})
    except Exception as e: print(f, "ERR", e)
P
