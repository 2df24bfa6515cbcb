#!/bin/bash
# r09: bounce-synchronous tracer with per-bounce re-binning (sweep_kernel, ARV2_SWEEP=<min rays>) against wave_kernel
OUT=gpurun_out/r09_sweep.log; : > $OUT
echo "== parity under ARV2_SWEEP=1 (every launch through sweep_kernel)" >> $OUT
ARV2_SWEEP=1 timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -5 >> $OUT
run() { echo "== $*" >> $OUT; WL=""; [[ "$1" == W=c4 ]] && WL="--workload c4"; env "$@" timeout 600 python bench.py $WL --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms', d['segments_per_step'])" >> $OUT 2>&1; }
run W=c4
run W=c4 ARV2_SWEEP=1
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=3 ARV2_SWEEP_DIR_BITS=4
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=3
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=0 ARV2_SWEEP_DIR_BITS=0
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_DIR_MAJOR=1
run W=c2x8M ARV2_BENCH_RAYS=8000000
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=3
run W=c2x1M ARV2_SWEEP=1
cat $OUT
