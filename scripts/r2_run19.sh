#!/bin/bash
# round-2 GPU session 19 (1 GPU): smoke() and the default bench line on the final tree
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke19.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke19.log
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"
tail -c 1500 gpurun_out/r2_bench19.json
