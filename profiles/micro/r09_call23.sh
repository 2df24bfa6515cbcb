#!/bin/bash
# r09 call 23: the successor let in only when a step enters its reduction (-DARV2_CONV_LATE_TRIGGER): period and timeline
OUT=gpurun_out/r09_conv_latetrigger.log; : > $OUT
python -m pytest tests/test_sweep_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
LABEL=late-trigger ARV2_LIB=$PWD/audiorenderingv2_b200/lib_lt/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=late-trigger+late-fft ARV2_CONV_LATE_FFT=1 ARV2_LIB=$PWD/audiorenderingv2_b200/lib_lt/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
ARV2_LIB=$PWD/audiorenderingv2_b200/lib_tr/libarv2.so python profiles/micro/conv_trace.py > gpurun_out/r09_conv_trace_latetrigger.log 2>&1
head -8 gpurun_out/r09_conv_trace_latetrigger.log >> $OUT; grep "^slot" gpurun_out/r09_conv_trace_latetrigger.log | head -12 >> $OUT
cat $OUT
