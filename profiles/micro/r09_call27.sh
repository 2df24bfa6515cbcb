#!/bin/bash
# r09 call 27 (gpurun --gpus 8): multi-GPU tests, then the bench at N = 8, 4, 2 (final tree of the round)
OUT=gpurun_out
python -m pytest tests/test_multigpu_gpu.py tests/test_cli_gpu.py -m gpu -x -q > $OUT/r09_tests_8gpu.log 2>&1; tail -2 $OUT/r09_tests_8gpu.log
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r09_bench_n$N.json 2> $OUT/r09_bench_n$N.err
  tail -c 200 $OUT/r09_bench_n$N.json
done
