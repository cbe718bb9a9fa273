#!/bin/bash
cd "$(dirname "$0")/.."
for b in 1 2 4 8; do python bench.py --steps 5 --warmup 3 --batch-per-gpu $b --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B=%s' % d['config']['global_batch'], 'pairs/s %.1f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], 'update us %.1f' % d['roofline']['us_per_launch_group'], 'upd TF %.1f' % d['roofline']['achieved'], 'lookup us %.1f frac %.3f' % (d['roofline_lookup']['us_per_launch'], d['roofline_lookup']['frac']))
"; done
