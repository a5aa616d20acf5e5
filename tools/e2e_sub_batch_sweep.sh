#!/bin/bash
# e2e leg of bench.py for several HostPipeline sub-batch sizes / policies, three runs each (zero-compressed transport)
for cfg in "16 1 0" "32 1 0" "64 1 0" "32 0 0" "32 1 1"; do set -- $cfg; for r in 1 2 3; do
python bench.py --no-cpu-baseline --e2e-sub-batch $1 --e2e-sub-all $2 --e2e-expand-main $3 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); e = d['e2e']
print('sub_batch', $1, 'all', $2, 'expand_on_main', $3, 'value', round(d['value']), 'e2e zc', round(e['value']), 'ms', round(e['ms_per_step'], 2), 'dense', round(e['dense_format']['value']), 'ratio', round(e['value'] / d['value'], 3))"
done; done
