#!/bin/bash
# `ncu --set full` captures of the kernels DESIGN.md quotes, one launch each, for the CURRENT build (after the plain commands
# have exited 0). Reports land in gpurun_out/<tag>_*.ncu-rep; read them here with `ncu -i ... --page details|raw|source --csv`.
#   gpurun -- 'bash tools/ncu_full_captures.sh r2'
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --attn-both"
F5="python tools/time_fit_shots.py --shot 5 --episodes 8 --algos 3 --distinct 2"
$B > $OUT/${TAG}_full_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/${TAG}_full_plain.log; exit 1; }
$F5 >> $OUT/${TAG}_full_plain.log 2>&1 || { echo "plain 5-shot failed"; tail -5 $OUT/${TAG}_full_plain.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_fit_resident -s 1 -c 1 -o $OUT/${TAG}_fit_resident -f $B > $OUT/${TAG}_full_ncu1.log 2>&1
$NCU -k regex:k_kproj_scores -s 1 -c 1 -o $OUT/${TAG}_kproj_tcgen05 -f $B > $OUT/${TAG}_full_ncu2.log 2>&1
$NCU -k regex:k_logits_iou_stream -s 1 -c 1 -o $OUT/${TAG}_logits_iou_stream -f $B > $OUT/${TAG}_full_ncu3.log 2>&1
$NCU -k regex:k_fit_l2 -s 1 -c 1 -o $OUT/${TAG}_fit_l2_5shot -f $F5 > $OUT/${TAG}_full_ncu4.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
