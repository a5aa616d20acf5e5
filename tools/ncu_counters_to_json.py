"""Turn gpurun_out/<tag>_kernel_counters.csv (tools/ncu_bench_kernels.sh) into profiles/<tag>_kernel_counters.json:
per kernel (launches averaged) the DRAM bytes, duration and pipe counters bench.py's roofline block cites, with the
fingerprint of the build they were captured on.   python tools/ncu_counters_to_json.py r2"""
import csv, io, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src = os.path.join(ROOT, "gpurun_out", f"{tag}_kernel_counters.csv")
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
per = {}
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1].split("<")[0].replace("void ", "").strip()
    key = (name, r["ID"])
    per.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    per[key]["_grid"] = r.get("Grid Size", "")
kernels = {}
for (name, _id), m in per.items():
    k = kernels.setdefault(name, {"launches": 0, "sum": {}})
    k["launches"] += 1
    k["grid"] = m.pop("_grid")
    for a, v in m.items():
        k["sum"][a] = k["sum"].get(a, 0.0) + v
out = {"tag": tag, "command": "python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --attn-both (ncu --clock-control none)",
       "build_fingerprint": open(os.path.join(ROOT, "gpurun_out", f"{tag}_build_fingerprint.txt")).read().strip(), "kernels": {}}
# per-kernel source fingerprints (few_shot_seg_cwt_b200.build.kernel_fingerprints): recorded only while the tree still IS the
# build the capture was taken on — a capture of one kernel stays valid when another kernel's source changes later
sys.path.insert(0, ROOT)
from few_shot_seg_cwt_b200 import build as B
if B.fingerprint() == out["build_fingerprint"]:
    out["kernel_source_fingerprints"] = B.kernel_fingerprints()
for name, k in kernels.items():
    n = k["launches"]
    avg = {a: v / n for a, v in k["sum"].items()}
    out["kernels"][name] = {
        "launches_captured": n, "grid": k["grid"],
        "duration_us": avg.get("gpu__time_duration.sum", 0.0) / 1e3,
        "dram_bytes_read": avg.get("dram__bytes_read.sum"), "dram_bytes_write": avg.get("dram__bytes_write.sum"),
        "l2_bytes": avg.get("lts__t_bytes.sum"),
        "lsu_wavefronts_shared": avg.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "shared_bank_conflict_wavefronts": avg.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "issue_active_pct": avg.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": avg.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_active_pct": avg.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "inst_executed_pipe_uniform": avg.get("sm__inst_executed_pipe_uniform.sum"),
        "inst_executed": avg.get("sm__inst_executed.sum"),
        "sm_cycles_elapsed_max": avg.get("sm__cycles_elapsed.max"),
    }
dst = os.path.join(ROOT, "profiles", f"{tag}_kernel_counters.json")
json.dump(out, open(dst, "w"), indent=1)
print(dst, {k: (round(v["duration_us"], 1), v["dram_bytes_read"]) for k, v in out["kernels"].items()})
