#!/bin/bash
# e2e leg of bench.py for several HostPipeline policies, two runs each (zero-compressed transport):
#   sub-batch size, sub-batch every batch (1) or only the first (0), expansion on the head's stream (1) or the copy stream (0), ramp
STEPS=${STEPS:-20}
for cfg in "64 0 1 8,8,16,32" "16 1 1 8" "16 1 1 " "32 1 1 8,8,16" "8 1 1 " "16 1 0 8" "32 1 1 8,8,8,8"; do set -- $cfg; for r in 1 2; do
python bench.py --no-cpu-baseline --steps $STEPS --e2e-sub-batch $1 --e2e-sub-all $2 --e2e-expand-main $3 --e2e-ramp "${4:-}" 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); e = d['e2e']
print('steps', d['steps'], 'sub_batch', $1, 'all', $2, 'expand_on_main', $3, 'ramp', '${4:-}', 'value', round(d['value']), 'e2e zc', round(e['value']), 'ms', round(e['ms_per_step'], 2), 'dense', round(e['dense_format']['value']), 'ratio', round(e['value'] / d['value'], 3))"
done; done
