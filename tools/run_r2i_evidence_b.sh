#!/bin/bash
# one gpurun call: counters + full ncu captures of the final build, and the 5-shot bench line with equal e2e pieces
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --shot 5 --episodes 36 > gpurun_out/r2i_bench_line_5shot.json 2> gpurun_out/r2i_bench5.err; tail -1 gpurun_out/r2i_bench5.err
bash tools/ncu_bench_kernels.sh r2i
bash tools/ncu_full_captures.sh r2i
for k in fit_resident kproj_tcgen05 logits_iou_stream fit_l2_5shot; do
  ncu -i gpurun_out/r2i_$k.ncu-rep --page details > gpurun_out/r2i_${k}_ncu_details.txt 2>/dev/null
done
ls -la gpurun_out | tail -30
