#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_exact.py -x -q 2>&1 | tail -8
timeout -s KILL 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv1_u8.py tests/test_gpu_trainnet.py -x -q 2>&1 | tail -8
timeout -s KILL 200 python tools/conv1_once.py > gpurun_out/r2_conv1_u8_c.jsonl 2> gpurun_out/r2_conv1_u8_c.err; cat gpurun_out/r2_conv1_u8_c.jsonl; tail -3 gpurun_out/r2_conv1_u8_c.err
timeout -s KILL 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; cut -c1-330 gpurun_out/r2_bench_e.json; tail -3 gpurun_out/r2_bench_e.err
