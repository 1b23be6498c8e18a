#!/bin/bash
# final single-GPU validation: tests, bench (ours + reference arm), workloads, sanitizer on the new kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -1 gpurun_out/r2_bench_final.json | cut -c1-400; tail -2 gpurun_out/r2_bench_final.err
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/r2_ref_final.json 2> gpurun_out/r2_ref_final.err; tail -1 gpurun_out/r2_ref_final.json | cut -c1-600; tail -2 gpurun_out/r2_ref_final.err
timeout -s KILL 600 python tools/bench_workloads.py all > gpurun_out/r2_workloads_final.jsonl 2> gpurun_out/r2_workloads_final.err; cut -c1-330 gpurun_out/r2_workloads_final.jsonl; tail -2 gpurun_out/r2_workloads_final.err
timeout -s KILL 900 compute-sanitizer --tool memcheck python tools/sanitize_kernels.py > gpurun_out/r2_sanitizer_memcheck_kernels_b.txt 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck_kernels_b.txt
timeout -s KILL 900 compute-sanitizer --tool racecheck python tools/sanitize_kernels.py > gpurun_out/r2_sanitizer_racecheck_kernels_b.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck_kernels_b.txt
