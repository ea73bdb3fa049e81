#!/usr/bin/env bash
# One-box evidence pass for a round: GPU tests, smoke, headline bench (both arms), ncu launch list of the bench
# command, ncu --set full captures of the dominant kernels, config-5 inference.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
TAG=${1:-r02f}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"
# launch list of the SAME bench command, eagerly issued so that every kernel is a separate launch
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also-c3 --no-graph > $O/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gru_fwd_kernel" -s 3 -c 1 -o $O/${TAG}_prof_gru_fwd \
  python tools/prof_gru_once.py > $O/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gru_bwd_kernel" -s 3 -c 1 -o $O/${TAG}_prof_gru_bwd \
  python tools/prof_gru_once.py > $O/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_tn_kernel" -s 3 -c 1 -o $O/${TAG}_prof_proj \
  python tools/prof_gru_once.py > $O/${TAG}_ncu_proj.log 2>&1; echo "ncu proj rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gru_cl_bwd_kernel" -s 1 -c 1 -o $O/${TAG}_prof_gru_cl_bwd \
  python tools/prof_cluster_once.py > $O/${TAG}_ncu_clbwd.log 2>&1; echo "ncu cluster bwd rc=$?"
python tools/probe_cluster.py > $O/${TAG}_probe_cluster.log 2>&1; echo "probe cluster rc=$?"
python tools/probe_proj_bf16.py > $O/${TAG}_probe_proj_bf16.log 2>&1; echo "probe proj bf16 rc=$?"
python tools/trace_overlap.py --out $O/${TAG}_timeline_c2_1gpu.csv > $O/${TAG}_timeline_c2_1gpu.log 2>&1; echo "timeline rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_bf16_kernel" -s 2 -c 1 -o $O/${TAG}_prof_proj_bf16 \
  python tools/probe_proj_bf16.py > $O/${TAG}_ncu_proj_bf16.log 2>&1; echo "ncu proj bf16 rc=$?"
python tools/sweep_phases.py --hidden 24 64 128 256 --steps 6 > $O/${TAG}_sweep_phases_c4_1gpu.jsonl 2>/dev/null; echo "c4 rc=$?"
python tools/bench_inference.py > $O/${TAG}_bench_inference_c5.json 2> $O/${TAG}_bench_inference_c5.err; echo "c5 rc=$?"; cat $O/${TAG}_bench_inference_c5.json | cut -c1-300
