#!/bin/bash
# one gpurun call: the fused logits/IoU kernel of the product build and of every variant under tools/variants
# (SRC=iou.cu python tools/build_variants.py 0:-DLS_TAPER=1 ...): parity tests, then tools/time_iou.py; log -> gpurun_out/$1
LOG=gpurun_out/${1:-iou_variants.txt}
{
for v in product $(ls tools/variants/libcwt_v*.so 2>/dev/null); do
  if [ "$v" = product ]; then unset CWT_LIB_PATH; else export CWT_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "logits_iou or episode_head_vs_golden or batch_intersection or argmax_ties" 2>&1 | tail -1
  timeout 120 python tools/time_iou.py 2>&1 | tail -1
done
} > $LOG 2>&1
cat $LOG
