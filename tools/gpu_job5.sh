#!/bin/bash
# 2-GPU job: product NCCL test, bench line at N=2, workloads at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/bench_workloads.py ppo > gpurun_out/r2_workloads_c_n1.jsonl 2> gpurun_out/r2_workloads_c_n1.err; cut -c1-700 gpurun_out/r2_workloads_c_n1.jsonl; tail -3 gpurun_out/r2_workloads_c_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_workloads.py all > gpurun_out/r2_workloads_c_n2.jsonl 2> gpurun_out/r2_workloads_c_n2.err; cut -c1-500 gpurun_out/r2_workloads_c_n2.jsonl; tail -4 gpurun_out/r2_workloads_c_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.json; tail -4 gpurun_out/r2_bench_n2.err
