set -u
O=gpurun_out
run() { # name, env..., args
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-also-c3 $EXTRA > $O/x15_$name.json 2> $O/x15_$name.err; echo "$name rc=$?"
}
EXTRA="" run base A=1
EXTRA="" run pair TIMEGAN_B200_BWD_PAIR=1
EXTRA="--bt 2" run bt2 A=1
EXTRA="" run forkbt2 TIMEGAN_B200_FORK_BT=2
python - <<'P'
import json
for f in ("base","pair","bt2","forkbt2"):
    try:
        d=json.loads(open(f"gpurun_out/x15_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
P
