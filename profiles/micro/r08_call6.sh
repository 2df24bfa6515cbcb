#!/bin/bash
# r08 call 6 (gpurun --gpus 8): the whole GPU suite on a multi-GPU box, then the driver's scaling sequence N = 1, 2, 4, 8
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/r08_tests_8gpu.log 2>&1; tail -4 $OUT/r08_tests_8gpu.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/r08_scale_ref.json 2> $OUT/r08_scale.err
python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/r08_scale_n1.json 2>> $OUT/r08_scale.err
for N in 2 4 8; do
  NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r08_scale_n$N.json 2> $OUT/r08_scale_n$N.err
  grep -c "NCCL INFO" $OUT/r08_scale_n$N.err; grep -m2 -i "nvls\|NCCL version" $OUT/r08_scale_n$N.err
done
audiorenderingv2_b200/lib/arv2_cli 2>&1 | head -2
ls -la $OUT/r08_scale*
