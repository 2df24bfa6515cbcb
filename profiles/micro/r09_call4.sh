#!/bin/bash
# r09 (4): sweeps as the default tracer of large launches: tests, C2 new-seed cost, C4 line, --set full capture of one mid-depth sweep
OUT=gpurun_out/r09_call4.log; : > $OUT
timeout 1200 python -m pytest tests/test_sweep_gpu.py tests/test_trace_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8 >> $OUT
python bench.py --skip-extras --no-cpu-baseline --steps 5 --warmup 3 2>gpurun_out/r09_err.log | tail -1 > gpurun_out/r09_c2_quick.json
python -c "import json; d=json.load(open('gpurun_out/r09_c2_quick.json')); print('c2', d['value'], d['ms_per_step'], 'first', d['first_render_ms'], 'new seed', d['new_seed_render_ms'], 'e2e', d['e2e']['value'])" >> $OUT
python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 3 --warmup 3 2>>gpurun_out/r09_err.log | tail -1 > gpurun_out/r09_c4_quick.json
python -c "import json; d=json.load(open('gpurun_out/r09_c4_quick.json')); print('c4', d['value'], d['ms_per_step'], 'first', d['first_render_ms'], 'new seed', d['new_seed_render_ms'])" >> $OUT
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 25 -c 1 -o gpurun_out/sweep_r09 -f python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_sweep_r09.log 2>&1
echo "sweep capture rc=$?" >> $OUT
ncu --set full --clock-control none --import-source on -k regex:wave_kernel -s 3 -c 1 -o gpurun_out/wave_c4_r09 -f env ARV2_SWEEP=0 python bench.py --workload c4 --skip-extras --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_wave_c4_r09.log 2>&1
echo "wave c4 capture rc=$?" >> $OUT
cat $OUT
