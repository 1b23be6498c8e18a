#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_conv1_u8.py -x -q 2>&1 | tail -15
timeout -s KILL 600 python -m pytest tests/test_gpu_trainnet.py tests/test_gpu_impala_host.py tests/test_gpu_engine.py -x -q 2>&1 | tail -6
timeout -s KILL 200 python tools/conv1_once.py > gpurun_out/r2_conv1_u8.jsonl 2> gpurun_out/r2_conv1_u8.err; cat gpurun_out/r2_conv1_u8.jsonl; tail -3 gpurun_out/r2_conv1_u8.err
timeout -s KILL 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; cut -c1-1500 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'shiftconv_fwd_kernel|wgrad_pair_kernel|gather_s2d' -c 12 -o gpurun_out/r2_conv1_u8 python tools/conv1_once.py --once 2>&1 | tail -2
