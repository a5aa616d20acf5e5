#!/bin/bash
# Re-capture the hardware counters the bench line's roofline block cites, for the CURRENT build, into
# gpurun_out/<tag>_kernel_counters.csv (+ the launch list); tools/ncu_counters_to_json.py turns the CSV into
# profiles/<tag>_kernel_counters.json, which bench.py reads (`roofline.traffic`, with the build fingerprint it was taken on).
#   gpurun -- 'bash tools/ncu_bench_kernels.sh r2'        (one GPU; never under torchrun)
# The plain command runs first and must exit 0 (B200_PROFILING.md); numbers printed under ncu are never bench values.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --attn-both"
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum,\
sm__cycles_elapsed.max,lts__t_bytes.sum,sm__inst_executed.sum
$CMD > $OUT/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_bench_launches_raw.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --metrics $METRICS --clock-control none \
    -k regex:'k_fit_resident|k_fit_l2|k_logits_iou_stream|k_rtf_stream|k_ftr_stream|k_kproj' -c 40 \
    --csv --log-file $OUT/${TAG}_kernel_counters.csv $CMD > $OUT/${TAG}_ncu_counters.log 2>&1
python -m few_shot_seg_cwt_b200.build --fingerprint > $OUT/${TAG}_build_fingerprint.txt
tail -3 $OUT/${TAG}_ncu_counters.log
wc -l $OUT/${TAG}_kernel_counters.csv $OUT/${TAG}_bench_launches_raw.csv
