#!/bin/bash
# developer script (one gpurun call): parity tests of the multi-shot L2-resident fit, then its timing at 2 / 3 / 5 shots for the product
# build and every variant library under tools/variants (SRC=fit_l2.cu python tools/build_variants.py 0:-DL2_TM=0); log -> gpurun_out/$1
LOG=gpurun_out/${1:-l2_check.txt}
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_l2 or 5shot or fit_full_size" 2>&1 | tail -3
for v in product $(ls tools/variants/libcwt_v*.so 2>/dev/null); do
  if [ "$v" = product ]; then unset CWT_LIB_PATH; else export CWT_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  timeout 200 python tools/time_fit_shots.py --shot 5 --episodes 32 2>&1 | tail -3
  timeout 200 python tools/time_fit_shots.py --shot 2 --episodes 32 --algos 3 2>&1 | tail -1
  timeout 200 python tools/time_fit_shots.py --shot 3 --episodes 32 --algos 3 2>&1 | tail -1
done
} > $LOG 2>&1
cat $LOG
