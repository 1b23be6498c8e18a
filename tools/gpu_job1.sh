#!/bin/bash
# round-2 first GPU job: parity of the new kernels, K1 timing matrix, reference arm, bench line (all bounded)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nproc
timeout 300 python -m pytest tests/test_gpu_vtrace.py -x -q 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_mlp.py -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_engines.py -q 2>&1 | tail -40
timeout 300 python tools/bench_k1.py > gpurun_out/r2_k1_matrix_a.jsonl 2> gpurun_out/r2_k1_matrix_a.err; cat gpurun_out/r2_k1_matrix_a.jsonl | cut -c1-400; tail -3 gpurun_out/r2_k1_matrix_a.err
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r2_ref_a.json 2> gpurun_out/r2_ref_a.err; tail -c 1200 gpurun_out/r2_ref_a.json; tail -5 gpurun_out/r2_ref_a.err
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 600 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
