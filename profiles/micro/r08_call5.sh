#!/bin/bash
# r08 call 5 (run with gpurun --gpus 2): multi-GPU tests, CLI, bench at N=1 and N=2
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/r08_tests_2gpu.log 2>&1; tail -4 $OUT/r08_tests_2gpu.log
python bench.py --steps 5 --warmup 3 > $OUT/r08_bench_n1.json 2> $OUT/r08_bench_n1.err; tail -c 600 $OUT/r08_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r08_bench_n2.json 2> $OUT/r08_bench_n2.err; tail -c 600 $OUT/r08_bench_n2.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r08_bench_ref.json 2>> $OUT/r08_bench_n1.err
ls -la $OUT/r08_bench*
