#!/bin/bash
# r09 call 21: two partitions per ring stage (half the bulk copies): tests, us per block, timeline
OUT=gpurun_out/r09_conv_pairs.log; : > $OUT
python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $OUT
LABEL=pairs python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=pairs+early-fft ARV2_CONV_EARLY_FFT=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=single-copies ARV2_LIB=$PWD/audiorenderingv2_b200/lib_sc/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
ARV2_LIB=$PWD/audiorenderingv2_b200/lib_tr/libarv2.so python profiles/micro/conv_trace.py > gpurun_out/r09_conv_trace_pairs.log 2>&1
head -8 gpurun_out/r09_conv_trace_pairs.log >> $OUT
cat $OUT
