#!/bin/bash
# Profiling recipe of this repo (B200_PROFILING.md): run under gpurun, one GPU.
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01'
# 1. plain run (must exit 0), 2. launch list with per-launch device time,
# 3. one --set full capture of the dominant kernels.  Reports land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"      # the default bench command (all legs: the launch list covers the first 600 launches)
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:wave_kernel|trace_kernel" -s 3 -c 1 -o gpurun_out/trace_$TAG -f $CMD > gpurun_out/ncu_trace_$TAG.log 2>&1
echo "trace capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_mask_kernel -s 5 -c 1 -o gpurun_out/rerender_$TAG -f $CMD > gpurun_out/ncu_rerender_$TAG.log 2>&1
echo "rerender (mask) capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rr_walk_kernel -s 5 -c 1 -o gpurun_out/rrwalk_$TAG -f $CMD > gpurun_out/ncu_rrwalk_$TAG.log 2>&1
echo "rerender (walk) capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_step_kernel -s 40 -c 1 -o gpurun_out/conv_$TAG -f $CMD > gpurun_out/ncu_conv_$TAG.log 2>&1
echo "conv capture rc=$?"
ls -la gpurun_out
