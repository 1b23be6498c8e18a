#!/bin/bash
# graphed learner update at small per-GPU batches: full suite, then A/B at the 8-GPU and 4-GPU shares on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
for E in 512 1024; do for G in 0 auto; do
echo "== bench envs=$E learn_graph=$G"; PARL_B200_LEARN_GRAPH=$G timeout -s KILL 300 python bench.py --envs $E --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2> gpurun_out/r2_bench_lg_${E}_$G.err | tail -1 > gpurun_out/r2_bench_lg_${E}_$G.json; python -c "
import sys,json
d=json.loads(open('gpurun_out/r2_bench_lg_${E}_$G.json').read().strip().splitlines()[-1]); print('envs $E graph $G: ms_per_step', round(d['ms_per_step'],3), 'value', int(d['value']), d['config']['learner_update'], 'launches', d['gpu_launches'])"
tail -2 gpurun_out/r2_bench_lg_${E}_$G.err
done; done
