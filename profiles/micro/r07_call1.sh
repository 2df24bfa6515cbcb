# r07 call 1: state check after restore + stack placement A/B + leaf-size A/B (one box)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r07_tests.log
VARIANTS="lib lib_ss8 lib_ss16 lib_ss24 lib_top lib_ss16top" bash profiles/micro/ab.sh > gpurun_out/r07_stack_ab.log 2>&1
CASES="leaf1 ARV2_LEAF_MAX=1
leaf2 ARV2_LEAF_MAX=2
leaf3 ARV2_LEAF_MAX=3
leaf4 ARV2_LEAF_MAX=4
leaf8 ARV2_LEAF_MAX=8" bash profiles/micro/env_ab.sh > gpurun_out/r07_leaf_ab.log 2>&1
( ARV2_LIB=$PWD/audiorenderingv2_b200/lib_ss16top/libarv2.so timeout 600 python -m pytest tests/test_trace_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r07_tests_ss16top.log
cat gpurun_out/r07_tests.log gpurun_out/r07_stack_ab.log gpurun_out/r07_leaf_ab.log gpurun_out/r07_tests_ss16top.log
