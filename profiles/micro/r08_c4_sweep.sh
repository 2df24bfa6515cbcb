#!/bin/bash
# C4 (1M-triangle hall, 8 bands, 10M rays): wave_kernel parameters, tuned on C2 in r05/r07
OUT=gpurun_out/r08_c4_sweep.log; : > $OUT
run() { echo "== $*" >> $OUT; env "$@" python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],2),'ms')" >> $OUT 2>&1; }
run A=1
run ARV2_WAVE_SEGMENTS=4
run ARV2_WAVE_SEGMENTS=16
run ARV2_WAVE_CAP=1024
run ARV2_WAVE_CAP=4096
run ARV2_WAVE_CAP=8192
run ARV2_WAVE_CAP=4096 ARV2_WAVE_SEGMENTS=16
run ARV2_NO_SORT=1
cat $OUT
