#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 200 python tools/conv_forms.py > gpurun_out/r2_conv_forms_a.jsonl 2> gpurun_out/r2_conv_forms_a.err; cat gpurun_out/r2_conv_forms_a.jsonl; tail -3 gpurun_out/r2_conv_forms_a.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'shiftconv_sfused' -c 10 -o gpurun_out/r2_conv_fused python tools/conv_forms.py --once 2>&1 | tail -2
