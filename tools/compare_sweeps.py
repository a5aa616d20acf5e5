"""Check that sweep lines (few_shot_seg_cwt_b200.sweep, one JSON line per file) taken at different world sizes hold the SAME
int64 intersection/union table:   python tools/compare_sweeps.py profiles/r2_sweep_10000_n1.json profiles/r2_sweep_10000_n2.json ..."""
import json, sys
lines = [json.loads([l for l in open(f).read().splitlines() if l.startswith('{"episodes"')][-1]) for f in sys.argv[1:]]
ref = lines[0]
ok = True
for f, d in zip(sys.argv[1:], lines):
    same = d["table_cls_I_U"] == ref["table_cls_I_U"] and d["table_fb_I_U"] == ref["table_fb_I_U"] and d["episodes"] == ref["episodes"]
    ok &= same
    print(f"{f}: world {d['world']} episodes {d['episodes']} sha256 {d['table_sha256'][:16]} mIoU {d['mIoU_adapted']:.10f} "
          f"{d['episodes_per_s_incl_generation']:.0f} episodes/s incl. generation  table identical to the first: {same}")
print("ALL TABLES IDENTICAL" if ok else "TABLES DIFFER")
sys.exit(0 if ok else 1)
