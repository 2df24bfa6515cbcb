#!/bin/bash
# r09 call 24: 16-CTA clusters per source for streams of <= 4 sources
OUT=gpurun_out/r09_conv_c16.log; : > $OUT
python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $OUT
LABEL=cluster16 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=cluster8 ARV2_CONV_CLUSTER16=0 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=cluster16-all ARV2_CONV_CLUSTER16=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=cluster16+late-fft ARV2_CONV_LATE_FFT=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
cat $OUT
