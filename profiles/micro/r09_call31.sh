#!/bin/bash
# r09 call 31 (gpurun --gpus 8): the bench at N = 8, 4, 2 with direction-tile shards (2^14 tiles): the round's final multi-GPU lines
OUT=gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r09_final_n$N.json 2> $OUT/r09_final_n$N.err
  grep '^{"metric"' $OUT/r09_final_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', round(d['value'],3), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), 'c3', round(d['c3_strong_grays_per_s'],2), 'c4', round(d['c4_grays_per_s'],2), 'conv', round(d['conv_us_per_block'],2), d['sharded_equals_single']['ok'], d['rank_kernel_ms'])"
done
