#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_impala_host.py tests/test_gpu_losses.py tests/test_gpu_mlp.py tests/test_gpu_replay.py tests/test_gpu_trainnet.py tests/test_gpu_vtrace.py -x -q 2>&1 | tail -40
for caps in "0 0" "64 84" "56 92" "48 100" "74 74"; do
set -- $caps
PARL_B200_ACTOR_SMS=$1 PARL_B200_LEARNER_SMS=$2 timeout -s KILL 300 python bench.py --envs 512 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('caps $1 $2 envs 512 ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"
done
