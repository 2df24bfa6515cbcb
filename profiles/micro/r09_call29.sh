#!/bin/bash
# r09 call 29 (gpurun --gpus 8): direction-tile shards: tests, then the bench at N = 8 and 2 against contiguous slices
OUT=gpurun_out
python -m pytest tests/test_trace_gpu.py tests/test_multigpu_gpu.py tests/test_cli_gpu.py -m gpu -x -q > $OUT/r09_tests_tiles.log 2>&1; tail -3 $OUT/r09_tests_tiles.log
for N in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r09_tiles_n$N.json 2> $OUT/r09_tiles_n$N.err
  grep '^{"metric"' $OUT/r09_tiles_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiles N=$N', round(d['value'],3), round(d['ms_per_step'],3), 'c3', round(d['c3_strong_grays_per_s'],2), 'c4', round(d['c4_grays_per_s'],2), d['sharded_equals_single'], d['rank_kernel_ms'])"
done
ARV2_SHARD_CONTIGUOUS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/r09_contig_n8.json 2> $OUT/r09_contig_n8.err
grep '^{"metric"' $OUT/r09_contig_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('contiguous N=8', round(d['value'],3), round(d['ms_per_step'],3), 'c3', round(d['c3_strong_grays_per_s'],2), 'c4', round(d['c4_grays_per_s'],2), d['rank_kernel_ms'])"
