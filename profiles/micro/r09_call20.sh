#!/bin/bash
# r09 call 20: segments per sweep with the kept bins (8^3 x 32^2): Grays/s and ncu lanes per instruction of mid-depth sweeps
OUT=gpurun_out/r09_sweep_seg.log; : > $OUT
run() { echo "== $*" >> $OUT; WL=""; [[ "$1" == W=c4 ]] && WL="--workload c4"; env "$@" timeout 600 python bench.py $WL --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms')" >> $OUT 2>&1; }
run W=c4 ARV2_SWEEP_SEGMENTS=2
run W=c4 ARV2_SWEEP_SEGMENTS=3
run W=c4 ARV2_SWEEP_SEGMENTS=2 ARV2_SWEEP_FIRST=6
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP_SEGMENTS=2
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP_SEGMENTS=3
run W=c2x30M ARV2_BENCH_RAYS=30000000 ARV2_SWEEP_SEGMENTS=2
run W=c2x30M ARV2_BENCH_RAYS=30000000 ARV2_SWEEP_SEGMENTS=3
for S in 2 3; do
  ARV2_SWEEP_SEGMENTS=$S ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum --clock-control none -k regex:sweep_kernel -s 16 -c 12 --csv --log-file gpurun_out/r09_lanes_s$S.csv python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
  echo "== ncu lanes per instruction, C4, sweeps 16..27 of the first render, $S segments per sweep" >> $OUT
  grep -v "^==" gpurun_out/r09_lanes_s$S.csv | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
print(' '.join(r['Metric Value'] for r in rows if 'thread_inst' in r['Metric Name']))" >> $OUT
done
cat $OUT
