set -u
O=gpurun_out
python tools/trace_overlap.py --out $O/r02_timeline_c2_1gpu.csv > $O/r02_timeline_c2_1gpu.log 2>&1; echo rc=$?; grep timeline $O/r02_timeline_c2_1gpu.log | head -20; tail -3 $O/r02_timeline_c2_1gpu.log
