#!/bin/bash
# one gpurun call: the round's last check — GPU test suite, smoke(), bench lines (default and 5 shots) of the final build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r2j_pytest.txt
cat gpurun_out/r2j_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2j_smoke.txt
python -m few_shot_seg_cwt_b200.build --fingerprint > gpurun_out/r2j_build_fingerprint.txt
timeout 400 python bench.py --steps 20 > gpurun_out/r2j_bench_line.json 2> gpurun_out/r2j_bench.err; tail -2 gpurun_out/r2j_bench.err
timeout 400 python bench.py --steps 10 --shot 5 --episodes 36 --no-cpu-baseline > gpurun_out/r2j_bench_line_5shot.json 2> gpurun_out/r2j_bench5.err; tail -2 gpurun_out/r2j_bench5.err
python - <<'PY'
import json
for f in ("r2j_bench_line", "r2j_bench_line_5shot"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d["roofline"]["bound"], round(d["roofline"]["frac"], 3),
              d["roofline"].get("traffic_source"), (d.get("parity_check") or {}).get("ok"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
