# A/B of library builds on one box: VARIANTS="lib lib_x ..." (directories under audiorenderingv2_b200/),
# optional RAYS_LIST="1000000 16000000"
for n in ${RAYS_LIST:-1000000}; do
for d in ${VARIANTS:-lib}; do
  ARV2_BENCH_RAYS=$n ARV2_LIB=$PWD/audiorenderingv2_b200/$d/libarv2.so timeout 300 python bench.py --steps 5 --warmup 3 --skip-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$d', $n, round(d['value'],3), 'Grays/s', round(d['roofline']['kernel_ms'],3), 'ms')"
done
done
