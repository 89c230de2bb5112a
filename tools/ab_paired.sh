#!/bin/bash
# A/B of the single-step kernel shapes (SPL_STEP_PAIRED=0/1), same box, alternating: lock-step microseconds from bench.py
for pz in 0 1 0 1 0 1; do
  SPL_STEP_PAIRED=$pz python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --skip-configs 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('paired=$pz lockstep us', round(d['lockstep']['us_per_lock_step'],2), 'mt19937 decks', round(d['lockstep']['bit_exact_decks']['us_per_lock_step'],2))
"
done
