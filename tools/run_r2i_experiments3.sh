#!/bin/bash
# developer script (one gpurun call): tensor-map staging of the resident fit (product) against the 2 560 bulk copies
# (python tools/build_variants.py "0:-DRES_TMA_STAGE=0"): fit parity tests, E = 64 timing, small batches
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_full_size or fit_resident or fit_batch_equals or fit_status or episode_head" 2>&1 | tail -2
} > gpurun_out/r2i_tma_stage.txt 2>&1
bash tools/run_ablation.sh r2i_tma_stage_timing.txt
cat gpurun_out/r2i_tma_stage_timing.txt >> gpurun_out/r2i_tma_stage.txt
{
timeout 100 python tools/ablate_resident.py --episodes 1 --tag "product E=1"
timeout 100 python tools/ablate_resident.py --episodes 4 --tag "product E=4"
CWT_LIB_PATH=$PWD/tools/variants/libcwt_v0_RES_TMA_STAGE0.so timeout 100 python tools/ablate_resident.py --episodes 1 --tag "bulk copies E=1"
} >> gpurun_out/r2i_tma_stage.txt 2>&1
cat gpurun_out/r2i_tma_stage.txt
