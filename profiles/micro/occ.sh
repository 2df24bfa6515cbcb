for c in 1 2 3 4 5 6 7; do
  ARV2_CTAS_PER_SM=$c timeout 300 python bench.py --steps 5 --warmup 3 --skip-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ctas/sm $c', round(d['value'],3), 'Grays/s', round(d['roofline']['kernel_ms'],3), 'ms')"
done
