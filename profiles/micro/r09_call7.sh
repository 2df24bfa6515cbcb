#!/bin/bash
# r09 call 7 (gpurun --gpus 2): NCCL binding order (multi-GPU tests FIRST in a fresh process), conv with the early forward FFT
OUT=gpurun_out
python -m pytest tests/test_multigpu_gpu.py tests/test_cli_gpu.py -m gpu -x -q > $OUT/r09_tests_2gpu.log 2>&1; tail -3 $OUT/r09_tests_2gpu.log
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > $OUT/r09_tests_conv.log 2>&1; tail -3 $OUT/r09_tests_conv.log
CUDA_VISIBLE_DEVICES=0 LABEL=early-fft python profiles/micro/conv_blocks.py > $OUT/r09_conv_ab.log 2>&1
CUDA_VISIBLE_DEVICES=0 LABEL=late-fft ARV2_CONV_LATE_FFT=1 python profiles/micro/conv_blocks.py >> $OUT/r09_conv_ab.log 2>&1
CUDA_VISIBLE_DEVICES=0 LABEL=early-fft python profiles/micro/conv_blocks.py >> $OUT/r09_conv_ab.log 2>&1
cat $OUT/r09_conv_ab.log
