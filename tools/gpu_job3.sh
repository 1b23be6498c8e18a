#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vtrace.py -x -q 2>&1 | tail -8
timeout 300 python tools/bench_k1.py > gpurun_out/r2_k1_matrix_b.jsonl 2> gpurun_out/r2_k1_matrix_b.err; cut -c1-330 gpurun_out/r2_k1_matrix_b.jsonl; tail -3 gpurun_out/r2_k1_matrix_b.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vtrace_loss_cta -c 2 -o gpurun_out/r2_k1_v6 python tools/k1_once.py 4096 0 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_engines.py -q 2>&1 | tail -12
timeout 600 python tools/bench_workloads.py ppo > gpurun_out/r2_workloads_b.jsonl 2> gpurun_out/r2_workloads_b.err; cut -c1-900 gpurun_out/r2_workloads_b.jsonl; tail -3 gpurun_out/r2_workloads_b.err
