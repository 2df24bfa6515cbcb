#!/bin/bash
# r09 call 11: timing experiment -- dependents triggered at kernel start (results invalid): what bounds the 6.4 us period?
OUT=gpurun_out/r09_conv_trigger.log; : > $OUT
LABEL=early-trigger ARV2_LIB=$PWD/audiorenderingv2_b200/lib_et/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=early-trigger+deep0 ARV2_CONV_DEEP_RING=0 ARV2_LIB=$PWD/audiorenderingv2_b200/lib_et/libarv2.so python profiles/micro/conv_blocks.py >> $OUT 2>&1
LABEL=no-pdl ARV2_CONV_NO_PDL=1 python profiles/micro/conv_blocks.py >> $OUT 2>&1
cat $OUT
