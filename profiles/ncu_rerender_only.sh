#!/bin/bash
TAG=${1:-x}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:rerender_kernel -s 5 -c 1 -o gpurun_out/rerender_$TAG -f $CMD > gpurun_out/ncu_rerender_$TAG.log 2>&1
echo "rerender capture rc=$?"
