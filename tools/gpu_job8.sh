#!/bin/bash
# N=8-share experiment on ONE GPU (512 envs = the per-GPU share of the 8-GPU run): CTA caps of the two streams
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { PARL_B200_ACTOR_SMS=$1 PARL_B200_LEARNER_SMS=$2 timeout 300 python bench.py --envs $3 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs',$3,'actor_sms','$1','learner_sms','$2','ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"; }
for cfg in "0 0" "74 74" "96 52" "52 96" "64 84" "110 38"; do run $cfg 512; done
for cfg in "0 0" "74 74" "96 52"; do run $cfg 1024; done
for cfg in "0 0" "100 100" "120 120"; do run $cfg 4096; done
timeout 300 python -m pytest tests/test_gpu_vtrace.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_k1.py > gpurun_out/r2_k1_matrix_c.jsonl 2> gpurun_out/r2_k1_matrix_c.err; cut -c1-330 gpurun_out/r2_k1_matrix_c.jsonl; tail -3 gpurun_out/r2_k1_matrix_c.err
