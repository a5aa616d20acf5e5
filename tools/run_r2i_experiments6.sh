#!/bin/bash
# developer script: halo gate in k_fit_l2 (product) against the variant without it (SRC=fit_l2.cu python tools/build_variants.py "0:-DL2_HALO_GATE=0")
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_l2 or 5shot or multi_shot" 2>&1 | tail -2
for v in product $(ls tools/variants/libcwt_v*.so 2>/dev/null); do
  if [ "$v" = product ]; then unset CWT_LIB_PATH; else export CWT_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  for s in 5 2 3; do timeout 200 python tools/time_fit_shots.py --shot $s --episodes 36 --algos 3 2>&1 | tail -1; done
done
} > gpurun_out/r2i_fit_l2_halo_gate.txt 2>&1
cat gpurun_out/r2i_fit_l2_halo_gate.txt
