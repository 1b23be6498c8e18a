#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ppo_launches.csv python tools/ppo_once.py 2>&1 | tail -2
python tools/ncu_summary.py launches gpurun_out/r2_ppo_launches.csv 2>&1 | head -40
timeout 300 python -m pytest tests/test_gpu_vtrace.py tests/test_gpu_engines.py -x -q 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_kernels.py > gpurun_out/r2_sanitizer_memcheck_kernels.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_memcheck_kernels.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_kernels.py > gpurun_out/r2_sanitizer_racecheck_kernels.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck_kernels.txt
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_kernels.py > gpurun_out/r2_sanitizer_synccheck_kernels.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_synccheck_kernels.txt
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_step.py 64 4 > gpurun_out/r2_sanitizer_memcheck_step.txt 2>&1; tail -8 gpurun_out/r2_sanitizer_memcheck_step.txt
