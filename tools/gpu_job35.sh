#!/bin/bash
# split-operand fp32 verification tests, host-contract tests, e2e after the set_weights / proxy changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest split + host"; timeout -s KILL 500 python -m pytest tests/test_gpu_split.py tests/test_gpu_impala_host.py -q 2>&1 | tail -12
echo "== bench e2e"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; tail -1 gpurun_out/r2_bench_j.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', int(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', int(e['value']), round(e['ms_per_step'],1), e['learner_thread_ms_per_step'], 'actor sample ms', e['actor_last_sample_ms'], 'k1', round(d['roofline_k1']['frac'],3), 'conv1', round(d['roofline']['frac'],3))"
tail -2 gpurun_out/r2_bench_j.err
