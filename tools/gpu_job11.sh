#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_conv1_u8.py -x -q 2>&1 | tail -8
timeout -s KILL 200 python tools/conv1_once.py > gpurun_out/r2_conv1_u8_b.jsonl 2> gpurun_out/r2_conv1_u8_b.err; cat gpurun_out/r2_conv1_u8_b.jsonl; tail -3 gpurun_out/r2_conv1_u8_b.err
timeout -s KILL 600 python -m pytest tests/test_gpu_trainnet.py tests/test_gpu_impala_host.py tests/test_gpu_engines.py tests/test_gpu_algorithms.py -x -q 2>&1 | tail -6
timeout -s KILL 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; cut -c1-400 gpurun_out/r2_bench_d.json; tail -3 gpurun_out/r2_bench_d.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'shiftconv_fwd_kernel|wgrad_pair_kernel' -c 8 -o gpurun_out/r2_conv1_u8 python tools/conv1_once.py --once 2>&1 | tail -2
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ppo_launches_b.csv python tools/ppo_once.py 2>&1 | tail -2
