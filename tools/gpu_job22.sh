#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for f in 0 1; do
echo "FUSE_STEP_GATHER=$f"; PARL_B200_FUSE_STEP_GATHER=$f timeout -s KILL 300 python tools/impala_phases.py 512 4096 2>&1 | tail -2
PARL_B200_FUSE_STEP_GATHER=$f timeout -s KILL 300 python bench.py --envs 512 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse $f envs 512 ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"
PARL_B200_FUSE_STEP_GATHER=$f timeout -s KILL 300 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse $f envs 4096 ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"
done
