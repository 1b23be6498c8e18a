#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout -s KILL 300 python tools/impala_phases.py 512 1024 4096 > gpurun_out/r2_phases.jsonl 2> gpurun_out/r2_phases.err; cat gpurun_out/r2_phases.jsonl; tail -3 gpurun_out/r2_phases.err
timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; cut -c1-3000 gpurun_out/r2_bench_f.json; tail -3 gpurun_out/r2_bench_f.err
