#!/bin/bash
for c in 1; do for b in 2 3 4; do
  CGB_FAST_CUBIC=$c CGB_FAST_MINB=$b python bench.py --steps 2 --warmup 3 --days-per-step 30 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('CUBIC=$c MINB=$b', 'value %.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'ms/step %.1f'%d['ms_per_step'], d['clocks'])"
done; done
