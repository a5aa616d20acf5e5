# developer script: parity tests of the fit + per-phase profile of the resident kernel variants
# (CFGS = list of TMEM:APPLIERS pairs; tile in tensor memory (1) or shared memory (0), applier warps 0 / 1 / 4)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_full_size or fit_resident or fit_batch_equals or fit_status" 2>&1 | tail -5
for cfg in ${CFGS:-1:0 1:4 0:1}; do
echo "=== TMEM=${cfg%%:*} APPLIERS=${cfg##*:}"
CWT_RESIDENT_TMEM=${cfg%%:*} CWT_RESIDENT_APPLIERS=${cfg##*:} timeout 300 python tools/prof_resident.py --episodes 64 2>&1 | tail -12
done
