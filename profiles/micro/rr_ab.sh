# re-render A/B over environment settings: each line of $CASES is "label ENV=.. ENV=.."
while read -r label envs; do
  [ -z "$label" ] && continue
  env $envs timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', 'rerender', round(d['rerender_ms'],4), 'ms device,', round(d['rerender_wall_ms'],4), 'wall, cache build', round(d['path_cache_build_ms'],2))"
done <<< "$CASES"
