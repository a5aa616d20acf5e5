#!/bin/bash
# one gpurun call: product timing + every variant under tools/variants (tools/build_variants.py); log -> gpurun_out/$1
LOG=gpurun_out/${1:-ablation.log}
mkdir -p gpurun_out
{
timeout 300 python tools/ablate_resident.py --save /tmp/w_ref.pt --tag product
for v in $(ls tools/variants/libcwt_v*.so 2>/dev/null); do
  CWT_LIB_PATH=$PWD/$v timeout 40 python tools/ablate_resident.py --ref /tmp/w_ref.pt --tag $(basename $v)
done
} > $LOG 2>&1
cat $LOG
