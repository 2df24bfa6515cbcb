# r07 call 3: full GPU suite on the new tree + PDL convolver; conv A/B; wave segments at 8M rays
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r07_tests3.log
VARIANTS="lib lib_c6 lib_c3" python profiles/micro/conv_ab.py > gpurun_out/r07_conv_ab.log 2>&1
ARV2_CONV_NO_PDL=1 VARIANTS="lib lib_c6" python profiles/micro/conv_ab.py 2>&1 | sed 's/^/nopdl /' >> gpurun_out/r07_conv_ab.log
for n in 1000000 8000000; do
CASES="seg4_$n ARV2_BENCH_RAYS=$n ARV2_WAVE_SEGMENTS=4
seg6_$n ARV2_BENCH_RAYS=$n ARV2_WAVE_SEGMENTS=6
seg8_$n ARV2_BENCH_RAYS=$n ARV2_WAVE_SEGMENTS=8
seg8cap1536_$n ARV2_BENCH_RAYS=$n ARV2_WAVE_SEGMENTS=8 ARV2_WAVE_CAP=1536" bash profiles/micro/env_ab.sh
done > gpurun_out/r07_seg_ab.log 2>&1
cat gpurun_out/r07_tests3.log gpurun_out/r07_conv_ab.log gpurun_out/r07_seg_ab.log
