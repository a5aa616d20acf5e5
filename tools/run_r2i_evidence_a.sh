#!/bin/bash
# one gpurun call: GPU test suite and the bench lines of the final build (default, 5 shots, tcgen05 attention)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2i_pytest_final.txt
cat gpurun_out/r2i_pytest_final.txt
python -m few_shot_seg_cwt_b200.build --fingerprint > gpurun_out/r2i_build_fingerprint.txt
timeout 600 python bench.py --steps 20 > gpurun_out/r2i_bench_line.json 2> gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_bench.err
timeout 600 python bench.py --steps 10 --shot 5 --episodes 36 > gpurun_out/r2i_bench_line_5shot.json 2> gpurun_out/r2i_bench5.err; tail -2 gpurun_out/r2i_bench5.err
timeout 600 python bench.py --steps 20 --attn-algo 1 --no-cpu-baseline > gpurun_out/r2i_bench_line_attn_tcgen05.json 2> gpurun_out/r2i_bench_attn.err; tail -2 gpurun_out/r2i_bench_attn.err
python - <<'PY'
import json
for f in ("r2i_bench_line", "r2i_bench_line_5shot", "r2i_bench_line_attn_tcgen05"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d["roofline"]["bound"], round(d["roofline"]["frac"], 3),
              d["roofline"].get("stages_ms"), (d.get("parity_check") or {}).get("ok"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
