#!/bin/bash
# r09 call 9: in-kernel stamps of a steady-state streaming step (-DARV2_CONV_TIMING), early and late forward FFT
OUT=gpurun_out/r09_conv_timing.log; : > $OUT
LABEL=early ARV2_LIB=$PWD/audiorenderingv2_b200/lib_ct/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=late ARV2_CONV_LATE_FFT=1 ARV2_LIB=$PWD/audiorenderingv2_b200/lib_ct/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
cat $OUT
