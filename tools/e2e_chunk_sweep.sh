#!/bin/bash
# e2e (pinned host in -> pinned host out) rate of both pipelines against the largest chunk size of Pipeline.run_host
for c in ${CHUNKS:-8 16 32 64}; do
  python bench.py --steps 10 --warmup 3 --no-cpu --chunk $c 2>/dev/null > gpurun_out/_c.json
  python -c "
import json; d=json.load(open('gpurun_out/_c.json')); g=d['config3_cnn_gf_x3']
print('chunk', $c, 'cnn_bf device', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '| cnn_gf_x3 device', round(g['value'],1), 'e2e', round(g['e2e']['value'],1), 'ms', round(g['e2e']['ms_per_step'],3))"
done
