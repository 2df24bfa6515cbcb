#!/bin/bash
# r09 call 40: bin geometry of the sweeps at 100 M rays (configs[2] on one GPU)
OUT=gpurun_out/r09_sweep_100m.log; : > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 600 python bench.py --skip-extras --no-cpu-baseline --steps 2 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms')" >> $OUT 2>&1; }
run ARV2_BENCH_RAYS=100000000
run ARV2_BENCH_RAYS=100000000 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=5
run ARV2_BENCH_RAYS=100000000 ARV2_SWEEP_CELL_BITS=3 ARV2_SWEEP_DIR_BITS=6
run ARV2_BENCH_RAYS=100000000 ARV2_SWEEP_CELL_BITS=4 ARV2_SWEEP_DIR_BITS=5 ARV2_SWEEP_SEGMENTS=1
cat $OUT
