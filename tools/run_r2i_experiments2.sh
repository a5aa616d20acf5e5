mkdir -p gpurun_out
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
