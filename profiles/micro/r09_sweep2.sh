#!/bin/bash
# r09 (2): tiled bin scan; per-kernel times of the sweeps (ncu launch list), bin geometry / segments per sweep
OUT=gpurun_out/r09_sweep2.log; : > $OUT
echo "== parity under ARV2_SWEEP=1" >> $OUT
ARV2_SWEEP=1 timeout 900 python -m pytest tests/test_trace_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $OUT
run() { echo "== $*" >> $OUT; WL=""; [[ "$1" == W=c4 ]] && WL="--workload c4"; env "$@" timeout 600 python bench.py $WL --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms', d['segments_per_step'])" >> $OUT 2>&1; }
run W=c4 ARV2_SWEEP=1
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=3
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=3
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=3 ARV2_SWEEP_SEGMENTS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=6 ARV2_SWEEP_DIR_BITS=2
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=5
run W=c4 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=3 ARV2_SWEEP_DIR_BITS=2
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1 ARV2_SWEEP_SEGMENTS=2
run W=c2x8M ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1 ARV2_SWEEP_CELL_BITS=5 ARV2_SWEEP_DIR_BITS=3
ARV2_SWEEP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/r09_launches_c4.csv python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r09_ncu_c4.log 2>&1
ARV2_BENCH_RAYS=8000000 ARV2_SWEEP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/r09_launches_c2x8m.csv python bench.py --skip-extras --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r09_ncu_c2.log 2>&1
cat $OUT
