mkdir -p gpurun_out
{
for T in 25 50 100 200 400; do timeout 100 python tools/ablate_resident.py --iters $T --tag "product T=$T"; done
} > gpurun_out/r2i_boundary_cost.txt 2>&1
cat gpurun_out/r2i_boundary_cost.txt
