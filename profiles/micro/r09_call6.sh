#!/bin/bash
# r09 call 6 (gpurun --gpus 8): multi-GPU tests on an 8-GPU box, then the bench at N = 8 and N = 2 (sweeps in configs[2] / [3])
OUT=gpurun_out
python -m pytest tests/test_multigpu_gpu.py tests/test_cli_gpu.py tests/test_sweep_gpu.py -m gpu -x -q > $OUT/r09_tests_8gpu.log 2>&1; tail -4 $OUT/r09_tests_8gpu.log
for N in 8 2; do
  NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r09_bench_n$N.json 2> $OUT/r09_bench_n$N.err
  grep -c "NCCL INFO" $OUT/r09_bench_n$N.err; tail -c 300 $OUT/r09_bench_n$N.json
done
