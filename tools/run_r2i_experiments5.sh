#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/e2e_policies.py > gpurun_out/r2i_e2e_policies.txt 2>&1
cat gpurun_out/r2i_e2e_policies.txt | tail -12
