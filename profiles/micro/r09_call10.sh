#!/bin/bash
# r09 call 10: deep ring (16 stages) for streams whose steps do not have to share SMs; conv tests; stamps
OUT=gpurun_out/r09_conv_deep.log; : > $OUT
python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $OUT
LABEL=deep+early python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=deep+late ARV2_CONV_LATE_FFT=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=shallow+early ARV2_CONV_DEEP_RING=0 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=all-deep ARV2_CONV_DEEP_RING=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=stamps ARV2_LIB=$PWD/audiorenderingv2_b200/lib_ct/libarv2.so python profiles/micro/conv_blocks.py 2>&1 | grep -A3 "sources 1:" | head -12 >> $OUT
cat $OUT
