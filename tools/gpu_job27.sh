#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; tail -1 gpurun_out/r2_bench_final2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', int(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), 'k1', round(d['roofline_k1']['frac'],3), 'step', round(d['roofline_step']['frac'],3), 'cpu', int(d['cpu_baseline']['value']), 'launches', d['gpu_launches'])"
tail -2 gpurun_out/r2_bench_final2.err
timeout -s KILL 300 python bench.py --envs 512 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs 512 ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']))"
