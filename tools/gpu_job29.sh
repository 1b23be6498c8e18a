#!/bin/bash
# round-2 session-2 job 1: K1 v8 parity + timing matrix + ncu + sanitizers + short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest vtrace"; timeout -s KILL 400 python -m pytest tests/test_gpu_vtrace.py -x -q 2>&1 | tail -15
echo "== bench_k1"; timeout -s KILL 200 python tools/bench_k1.py > gpurun_out/r2_k1_matrix_e.jsonl 2> gpurun_out/r2_k1_matrix_e.err; cat gpurun_out/r2_k1_matrix_e.jsonl | cut -c1-400; tail -3 gpurun_out/r2_k1_matrix_e.err
echo "== ncu"; timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:vtrace -o gpurun_out/r2_k1_v8b -f python tools/k1_once.py 4096 0 2>&1 | tail -3
timeout -s KILL 120 python tools/ncu_summary.py kernel gpurun_out/r2_k1_v8b.ncu-rep > gpurun_out/r2_k1_v8b_summary.txt 2>/dev/null; head -40 gpurun_out/r2_k1_v8b_summary.txt | cut -c1-160
echo "== memcheck"; timeout -s KILL 200 compute-sanitizer --tool memcheck python tools/k1_once.py 64 0 2>&1 | tail -4
echo "== racecheck"; timeout -s KILL 200 compute-sanitizer --tool racecheck python tools/k1_once.py 64 0 2>&1 | tail -4
echo "== bench short"; timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -1 gpurun_out/r2_bench_g.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'k1', d['roofline_k1'])"
tail -3 gpurun_out/r2_bench_g.err
