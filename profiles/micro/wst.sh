# scheduling statistics of wave_kernel (library built with -DARV2_WAVESTAT into lib_wst):
# counters: [8] tasks [9] rays taken [10] lost claims [11] idle polls
export ARV2_LIB=$PWD/audiorenderingv2_b200/lib_wst/libarv2.so ARV2_PRINT_STATS=1
for cfg in "X=1" "ARV2_WAVE_SEGMENTS=2" "ARV2_WAVE_CAP=4096"; do
echo "== $cfg"; env $cfg timeout 120 python bench.py --steps 1 --warmup 3 --skip-extras --no-cpu-baseline 2>&1 | grep -E "^[0-9]+ [0-9]+ " | tail -1
done
