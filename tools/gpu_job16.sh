#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 120 python tools/gemm_cluster.py --small 2>&1 | tail -4
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_exact.py -x -q 2>&1 | tail -5
timeout -s KILL 200 python tools/gemm_cluster.py > gpurun_out/r2_gemm_cluster.jsonl 2> gpurun_out/r2_gemm_cluster.err; cat gpurun_out/r2_gemm_cluster.jsonl; tail -3 gpurun_out/r2_gemm_cluster.err
