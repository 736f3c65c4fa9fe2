#!/bin/bash
# round-2 GPU session 15 (1 GPU): validation of the final tree -- full GPU suite, smoke(), launch list, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t15.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t15.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke15.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke15.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_B256_v6.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/r2_ncu_launches15.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_B256_v6.csv > gpurun_out/r2_launch_shares_step_B256_v6.txt 2>&1
head -3 gpurun_out/r2_launch_shares_step_B256_v6.txt
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"
tail -c 3000 gpurun_out/r2_bench15.json
S=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/r2_bench15_ref.json 2> gpurun_out/r2_bench15_ref.err; echo "reference arm rc=$? in $(( $(date +%s) - S )) s"
cat gpurun_out/r2_bench15_ref.json
