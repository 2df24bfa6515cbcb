#!/bin/bash
# r09 (3): first sweep of 8 segments, segments per sweep, bin geometry, crossover with the launch size, lanes per node step
OUT=gpurun_out/r09_sweep3.log; : > $OUT
run() { echo "== $*" >> $OUT; WL=""; [[ "$1" == W=c4 ]] && WL="--workload c4"; env "$@" timeout 600 python bench.py $WL --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms', d['segments_per_step'])" >> $OUT 2>&1; }
run W=c4 ARV2_SWEEP=1
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_FIRST=4
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_FIRST=12
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=4
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=6
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=4
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=5
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=3 ARV2_SWEEP_DIR_BITS=5
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=2 ARV2_SWEEP_DIR_BITS=6
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=5
run W=c2x4M ARV2_BENCH_RAYS=4000000
run W=c2x4M ARV2_BENCH_RAYS=4000000 ARV2_SWEEP=1
run W=c2x2M ARV2_BENCH_RAYS=2000000
run W=c2x2M ARV2_BENCH_RAYS=2000000 ARV2_SWEEP=1
run W=c2x1M ARV2_SWEEP=1
run W=c2x30M ARV2_BENCH_RAYS=30000000
run W=c2x30M ARV2_BENCH_RAYS=30000000 ARV2_SWEEP=1
echo "== traversal tallies, C4 10M rays: wave / sweeps" >> $OUT
ARV2_LIB=$PWD/audiorenderingv2_b200/lib/libarv2_stats.so ARV2_STATS_RAYS=10000000 python bench.py --stats-pass --workload c4 2>/dev/null | tail -1 >> $OUT
ARV2_SWEEP=1 ARV2_LIB=$PWD/audiorenderingv2_b200/lib/libarv2_stats.so ARV2_STATS_RAYS=10000000 python bench.py --stats-pass --workload c4 2>/dev/null | tail -1 >> $OUT
ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=1 ARV2_SWEEP_FIRST=1 ARV2_LIB=$PWD/audiorenderingv2_b200/lib/libarv2_stats.so ARV2_STATS_RAYS=10000000 python bench.py --stats-pass --workload c4 2>/dev/null | tail -1 >> $OUT
cat $OUT
