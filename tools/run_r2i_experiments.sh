#!/bin/bash
# developer script (one gpurun call, round 2 last session): GPU test suite, accumulator-stride variants of the resident fit
# (python tools/build_variants.py "0:-DRES_ACC_SHIFT=2" ...), inner_loop variant timings, 5-shot plan with 4 tiles per CTA x 3 groups
mkdir -p gpurun_out
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
