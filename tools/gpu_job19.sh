#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_trainnet.py -x -q 2>&1 | tail -12
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 600 --csv --log-file gpurun_out/r2_launches_512.csv python bench.py --envs 512 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > gpurun_out/r2_launches_512.out 2>&1; wc -l gpurun_out/r2_launches_512.csv; tail -5 gpurun_out/r2_launches_512.out | cut -c1-300
