#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_impala_host.py -q 2>&1 | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vtrace_loss -s 2 -c 4 -o gpurun_out/r2_k1_v4v5 python tools/k1_once.py 4096 1,0 2>&1 | tail -5
timeout 600 python tools/bench_workloads.py all > gpurun_out/r2_workloads_a.jsonl 2> gpurun_out/r2_workloads_a.err; cut -c1-900 gpurun_out/r2_workloads_a.jsonl; tail -5 gpurun_out/r2_workloads_a.err
