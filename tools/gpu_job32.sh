#!/bin/bash
# e2e: slab learner + column-group actor, phase breakdown; A/B over the number of actor groups
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest host"; timeout -s KILL 500 python -m pytest tests/test_gpu_impala_host.py -x -q 2>&1 | tail -8
for G in 4 1 2; do
echo "== bench e2e groups=$G"; PARL_B200_ACTOR_GROUPS=$G timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_e2e_g$G.json 2> gpurun_out/r2_bench_e2e_g$G.err; tail -1 gpurun_out/r2_bench_e2e_g$G.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', int(d['value']), 'e2e', int(e['value']), round(e['ms_per_step'],1), e['learner_thread_ms_per_step'], 'actor sample ms', e['actor_last_sample_ms'], e['actor_groups'])"
tail -2 gpurun_out/r2_bench_e2e_g$G.err
done
