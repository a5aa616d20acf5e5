#!/bin/bash
# developer script (one gpurun call): parity tests of the fit, then the product resident kernel and every variant library under
# tools/variants (python tools/build_variants.py 0:-DRES_POLL_DELAY=800 0x04 ...) timed at E = 64 x 200 steps; log -> gpurun_out/$1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_full_size or fit_resident or fit_batch_equals or fit_status" 2>&1 | tail -3
bash tools/run_ablation.sh ${1:-tmem_check.txt}
