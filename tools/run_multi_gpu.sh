#!/bin/bash
# one gpurun --gpus N call: bench.py and the 10 000-episode sweep on N GPUs (torchrun for N > 1); outputs gpurun_out/<tag>_*_Ngpu.json
N=${1:-2}; TAG=${2:-r2g}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then RUN="python"; fi
$RUN bench.py --gpus $N --no-cpu-baseline > gpurun_out/${TAG}_bench_line_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 300 gpurun_out/${TAG}_bench_line_${N}gpu.json; echo
if [ "$N" = "1" ]; then SW="python -m few_shot_seg_cwt_b200.sweep"; else SW="$RUN -m few_shot_seg_cwt_b200.sweep"; fi
$SW --episodes 10000 > gpurun_out/${TAG}_sweep_10000_n${N}.json 2> gpurun_out/${TAG}_sweep_n${N}.err
head -c 400 gpurun_out/${TAG}_sweep_10000_n${N}.json; echo
