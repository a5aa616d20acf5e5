# developer script: fit parity tests + timing of the product resident kernel (and the libraries under tools/variants)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fit_full_size or fit_resident or fit_batch_equals or fit_status" 2>&1 | tail -3
bash tools/run_ablation.sh ${1:-tmem_check.txt}
