#!/bin/bash
# r09 call 8: rank 0's share of the old partitions now that its forward FFT runs before the wait (ARV2_CONV_FFTCOST builds)
OUT=gpurun_out/r09_conv_fftcost.log; : > $OUT
for c in "" _fc16 _fc24 _fc32 _fc40; do
  LABEL="fftcost${c:-_8}" ARV2_LIB=$PWD/audiorenderingv2_b200/lib$c/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
done
cat $OUT
