#!/bin/bash
# GPU box: sweep the k_anyhit tuning knobs on the 1080p bench frame
for cfg in "24 8 8" "24 16 8" "24 4 8" "8 8 8" "64 8 8" "24 8 4" "24 8 12" "24 12 16" "48 12 12"; do
  set -- $cfg
  RT580_AH_STEPS=$1 RT580_AH_MIN_SEARCH=$2 RT580_AH_BLOCKS_PER_SM=$3 python bench.py --steps 2 --warmup 1 --width 1920 --height 1080 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('steps $1 min $2 bpsm $3 ->', round(d['value']), 'Mrays/s  ao_kernel_ms', round(d['roofline']['kernel_ms'],2), 'struct', round(d['phases_ms']['structure'],2))"
done
