#!/bin/bash
# r09 call 35: ray order by counting sort (top bits of the direction key) against the four-pass radix sort
OUT=gpurun_out/r09_order.log; : > $OUT
python -m pytest tests/test_trace_gpu.py tests/test_sweep_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
run() { echo "== $*" >> $OUT; WL=""; [[ "$1" == W=c4 ]] && WL="--workload c4"; env "$@" timeout 600 python bench.py $WL --skip-extras --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],3),'ms', 'first', round(d['first_render_ms'],2), 'new seed', round(d['new_seed_render_ms'],2))" >> $OUT 2>&1; }
run W=c2 A=counting
run W=c2 ARV2_ORDER_RADIX=1
run W=c2 A=counting


cat $OUT
