#!/bin/bash
# ncu launch list of the default bench command (device-resident part): shares of the step's kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1000 -c 1400 --csv --log-file gpurun_out/r2_launches_s2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_s2.log 2>&1; tail -2 gpurun_out/r2_launches_s2.log | cut -c1-300
python tools/ncu_summary.py launches gpurun_out/r2_launches_s2.csv > gpurun_out/r2_launches_s2.txt 2>&1; head -40 gpurun_out/r2_launches_s2.txt | cut -c1-170
