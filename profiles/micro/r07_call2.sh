# r07 call 2: SAH leaf test default, leaf pair loads, speculative traversal, wave parameters with the new tree
mkdir -p gpurun_out
VARIANTS="lib lib_pair lib_spec lib_specpair" bash profiles/micro/ab.sh > gpurun_out/r07_leafvar_ab.log 2>&1
CASES="oldtree ARV2_SAH_CT=0 ARV2_LEAF_MAX=4
ct1 ARV2_SAH_CT=1
ct1_leaf8 ARV2_SAH_CT=1 ARV2_LEAF_MAX=8
ct2_leaf8 ARV2_SAH_CT=2 ARV2_LEAF_MAX=8
seg6 ARV2_WAVE_SEGMENTS=6
seg8 ARV2_WAVE_SEGMENTS=8
cap1536 ARV2_WAVE_CAP=1536
cap3072 ARV2_WAVE_CAP=3072" bash profiles/micro/env_ab.sh > gpurun_out/r07_env_ab.log 2>&1
for e in "ARV2_SAH_CT=0" "ARV2_SAH_CT=1"; do
  env $e timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --skip-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 $e', round(d['value'],3), 'Grays/s', round(d['roofline']['kernel_ms'],3), 'ms')"
done > gpurun_out/r07_c4_ab.log 2>&1
( timeout 600 python -m pytest tests/test_trace_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r07_tests2.log
cat gpurun_out/r07_leafvar_ab.log gpurun_out/r07_env_ab.log gpurun_out/r07_c4_ab.log gpurun_out/r07_tests2.log
