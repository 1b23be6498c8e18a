#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_mlp.py tests/test_gpu_impala_host.py -x -q 2>&1 | tail -5
timeout 900 python tools/bench_workloads.py all > gpurun_out/r2_workloads_d.jsonl 2> gpurun_out/r2_workloads_d.err; cut -c1-700 gpurun_out/r2_workloads_d.jsonl; tail -3 gpurun_out/r2_workloads_d.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_mlp -s 1 -c 1 -o gpurun_out/r2_rollout_mlp python tools/rollout_once.py 64 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py losses > gpurun_out/r2_kernels_losses.jsonl 2> gpurun_out/r2_kernels_losses.err; cat gpurun_out/r2_kernels_losses.jsonl | cut -c1-300
timeout 600 ncu --set full --clock-control none -k regex:'flat_categorical|ppo_gaussian|gae_scan|td_loss|per_sample|per_write|replay_gather|gather_rows|mlp_fwd|mlp_bwd' -c 24 -o gpurun_out/r2_k2k4 python tools/bench_kernels.py losses 2>&1 | tail -2
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 2500 gpurun_out/r2_bench_b.json; tail -6 gpurun_out/r2_bench_b.err
timeout 300 python examples/impala_train.py --seconds 12 --env_num 1024 2>&1 | tail -4
