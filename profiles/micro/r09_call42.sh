#!/bin/bash
# r09 call 42: wave_kernel takes the unstarted rays per SM in chunks of consecutive rays of the direction order
OUT=gpurun_out/r09_wave_chunk.log; : > $OUT
ARV2_WAVE_CHUNK=1024 python -m pytest tests/test_trace_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 600 python bench.py --skip-extras --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r09_err.log | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print(round(d['value'],4),'Grays/s', round(d['ms_per_step'],3),'ms')" >> $OUT 2>&1; }
run A=32
run ARV2_WAVE_CHUNK=128
run ARV2_WAVE_CHUNK=512
run ARV2_WAVE_CHUNK=2048
run ARV2_WAVE_CHUNK=8192
run A=32
cat $OUT
