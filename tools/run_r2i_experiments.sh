#!/bin/bash
# Developer script: the experiment batches of the last session of round 2, one gpurun call each (outputs -> gpurun_out/r2i_*,
# the kept ones are under profiles/).   gpurun -- "bash tools/run_r2i_experiments.sh <batch>"
mkdir -p gpurun_out
case "${1:-1}" in
1)
# developer script (one gpurun call, round 2 last session): GPU test suite, accumulator-stride variants of the resident fit
# (python tools/build_variants.py "0:-DRES_ACC_SHIFT=2" ...), inner_loop variant timings, 5-shot plan with 4 tiles per CTA x 3 groups
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2i_pytest.txt
cat gpurun_out/r2i_pytest.txt
bash tools/run_ablation.sh r2i_acc_stride.txt
timeout 300 python tools/time_inner_loop_variants.py > gpurun_out/r2i_inner_loop_variants.json 2> gpurun_out/r2i_inner_loop_variants.err
tail -3 gpurun_out/r2i_inner_loop_variants.err
{
timeout 200 python tools/time_fit_shots.py --shot 5 --episodes 36 --algos 3 2>&1 | tail -1
CWT_FIT_L2_NT=4 CWT_FIT_L2_MB=150 timeout 200 python tools/time_fit_shots.py --shot 5 --episodes 36 --algos 3 2>&1 | tail -1
CWT_FIT_L2_NT=4 CWT_FIT_L2_MB=100 timeout 200 python tools/time_fit_shots.py --shot 5 --episodes 36 --algos 3 2>&1 | tail -1
} > gpurun_out/r2i_fit_l2_5shot_plans.txt 2>&1
cat gpurun_out/r2i_fit_l2_5shot_plans.txt
;;
2)
timeout 300 python tools/overlap_post_stage.py > gpurun_out/r2i_overlap_priority.txt 2>&1
cat gpurun_out/r2i_overlap_priority.txt | tail -9
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_l2 or 5shot" 2>&1 | tail -2
timeout 200 python tools/time_fit_shots.py --shot 5 --episodes 36 2>&1 | tail -3
timeout 200 python tools/time_fit_shots.py --shot 7 --episodes 16 --algos 3 2>&1 | tail -1
CWT_FIT_L2_NT=2 timeout 200 python tools/time_fit_shots.py --shot 7 --episodes 16 --algos 3 2>&1 | tail -1
timeout 200 python tools/time_fit_shots.py --shot 8 --episodes 16 --algos 3 2>&1 | tail -1
CWT_FIT_L2_NT=2 timeout 200 python tools/time_fit_shots.py --shot 8 --episodes 16 --algos 3 2>&1 | tail -1
} > gpurun_out/r2i_fit_l2_plans_after.txt 2>&1
cat gpurun_out/r2i_fit_l2_plans_after.txt
;;
3)
# developer script (one gpurun call): tensor-map staging of the resident fit (product) against the 2 560 bulk copies
# (python tools/build_variants.py "0:-DRES_TMA_STAGE=0"): fit parity tests, E = 64 timing, small batches
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
;;
4)
{
for T in 25 50 100 200 400; do timeout 100 python tools/ablate_resident.py --iters $T --tag "product T=$T"; done
} > gpurun_out/r2i_boundary_cost.txt 2>&1
cat gpurun_out/r2i_boundary_cost.txt
;;
5)
timeout 400 python tools/e2e_policies.py > gpurun_out/r2i_e2e_policies.txt 2>&1
cat gpurun_out/r2i_e2e_policies.txt | tail -12
;;
6)
# developer script: halo gate in k_fit_l2 (product) against the variant without it (SRC=fit_l2.cu python tools/build_variants.py "0:-DL2_HALO_GATE=0")
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_l2 or 5shot or multi_shot" 2>&1 | tail -2
for v in product $(ls tools/variants/libcwt_v*.so 2>/dev/null); do
  if [ "$v" = product ]; then unset CWT_LIB_PATH; else export CWT_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  for s in 5 2 3; do timeout 200 python tools/time_fit_shots.py --shot $s --episodes 36 --algos 3 2>&1 | tail -1; done
done
} > gpurun_out/r2i_fit_l2_halo_gate.txt 2>&1
cat gpurun_out/r2i_fit_l2_halo_gate.txt
;;
*) echo "usage: $0 1..6"; exit 2;;
esac
