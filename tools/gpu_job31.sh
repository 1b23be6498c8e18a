#!/bin/bash
# K1 v8 with the parallel tail + slab-pipelined host learner (e2e)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest vtrace + host"; timeout -s KILL 500 python -m pytest tests/test_gpu_vtrace.py tests/test_gpu_impala_host.py -x -q 2>&1 | tail -15
echo "== bench_k1"; timeout -s KILL 200 python tools/bench_k1.py > gpurun_out/r2_k1_matrix_g.jsonl 2> gpurun_out/r2_k1_matrix_g.err; cat gpurun_out/r2_k1_matrix_g.jsonl | cut -c1-330; tail -3 gpurun_out/r2_k1_matrix_g.err
echo "== ncu"; timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:vtrace -o gpurun_out/r2_k1_v8d -f python tools/k1_once.py 4096 0 2>&1 | tail -2
timeout -s KILL 120 python tools/ncu_summary.py kernel gpurun_out/r2_k1_v8d.ncu-rep > gpurun_out/r2_k1_v8d_summary.txt 2>/dev/null; head -22 gpurun_out/r2_k1_v8d_summary.txt | cut -c1-160
echo "== bench with e2e"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; tail -1 gpurun_out/r2_bench_i.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'k1', round(d['roofline_k1']['frac'],3), d['roofline_k1']['us_per_launch'], 'e2e', int(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['copy_bandwidth_gbs'], d['e2e']['last_losses'])"
tail -3 gpurun_out/r2_bench_i.err
