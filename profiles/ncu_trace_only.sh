#!/bin/bash
# one --set full capture of the trace kernel only (short: no extras)
TAG=${1:-x}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extras"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:wave_kernel|trace_kernel" -s 3 -c 1 -o gpurun_out/trace_$TAG -f $CMD > gpurun_out/ncu_trace_$TAG.log 2>&1
echo "trace capture rc=$?"
