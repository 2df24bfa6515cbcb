# A/B over environment settings on one box: each line of $CASES is "label ENV=.. ENV=.."
while read -r label envs; do
  [ -z "$label" ] && continue
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --skip-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['value'],3), 'Grays/s', round(d['roofline']['kernel_ms'],3), 'ms', d['segments_per_step'])"
done <<< "$CASES"
