for c in 16 32 64; do
  python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --chunk $c 2>/dev/null > gpurun_out/_c.json
  python -c "
import json; d=json.load(open('gpurun_out/_c.json')); print('chunk', $c, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))"
done
