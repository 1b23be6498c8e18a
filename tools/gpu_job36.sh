#!/bin/bash
# final validation of the session: full GPU suite, smoke, sanitizers on the final K1, default bench, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for tool in memcheck racecheck synccheck; do
echo "== $tool K1"; timeout -s KILL 200 compute-sanitizer --tool $tool python tools/k1_once.py 64 0,4,9 > gpurun_out/r2_sanitizer_${tool}_k1_v8.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_${tool}_k1_v8.txt
done
echo "== bench default"; timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_final_s2.json 2> gpurun_out/r2_bench_final_s2.err; tail -1 gpurun_out/r2_bench_final_s2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', int(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', int(e['value']), round(e['ms_per_step'],1), e['learner_thread_ms_per_step'], 'actor', e['actor_last_sample_ms'], 'bidir', e['copy_bandwidth_bidirectional_all_ranks_gbs'], int(e['pcie_ceiling_bidirectional_env_steps_per_s'] or 0), 'k1', round(d['roofline_k1']['frac'],3), 'conv1', round(d['roofline']['frac'],3), 'cpu', int(d['cpu_baseline']['value']), 'launches', d['gpu_launches'], d['clocks'])"
tail -2 gpurun_out/r2_bench_final_s2.err
echo "== launch list"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_s2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1; python tools/ncu_summary.py launches gpurun_out/r2_launches_s2.csv > gpurun_out/r2_launches_s2.txt 2>&1; head -24 gpurun_out/r2_launches_s2.txt | cut -c1-150
echo "== ncu gemm (fc shapes)"; timeout -s KILL 300 ncu --set full --clock-control none -k regex:gemm_bf16_tn_kernel -c 8 -o gpurun_out/r2_gemm_fc -f python tools/gemm_once.py > /dev/null 2>&1; timeout -s KILL 120 python tools/ncu_summary.py kernel gpurun_out/r2_gemm_fc.ncu-rep > gpurun_out/r2_gemm_fc_summary.txt 2>/dev/null; grep -E "Kernel Name|grid_size|time_duration|dram_throughput|tensor_cycles|wavefronts_mem_shared.sum" gpurun_out/r2_gemm_fc_summary.txt | cut -c1-150 | head -40
