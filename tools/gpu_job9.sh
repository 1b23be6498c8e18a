#!/bin/bash
# 2-GPU: product NCCL test (diagnostics) ; then single-GPU items on GPU 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -30
export CUDA_VISIBLE_DEVICES=0
timeout 600 python -m pytest tests/test_gpu_engines.py -x -q 2>&1 | tail -4
timeout 600 python tools/bench_workloads.py a2c ppo > gpurun_out/r2_workloads_e.jsonl 2> gpurun_out/r2_workloads_e.err; cut -c1-620 gpurun_out/r2_workloads_e.jsonl; tail -3 gpurun_out/r2_workloads_e.err
timeout 600 ncu --set full --clock-control none -k regex:'flat_categorical|ppo_gaussian|gae_scan|td_loss|per_sample|per_write|replay_gather|mlp_fwd|mlp_bwd|mlp_grad|adv_stats' -c 40 -o gpurun_out/r2_k2k4 python tools/kernels_once.py 2>&1 | tail -2
timeout 300 python bench.py --envs 512 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs 512 default caps ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"
