"""Summarise the nvcc -Xptxas -v logs written by few_shot_seg_cwt_b200.build (registers / spills / smem)."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "few_shot_seg_cwt_b200", "lib", "obj")
rows = []
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt):
        rows.append((os.path.basename(log).split(".")[0], m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), m.group(7) or "0"))
names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.splitlines() if rows else []
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("cwt::", "")
    if flt in n:
        print(f"{r[0]:14s} regs={r[2]:3d} stack={r[3]:4d} spill={r[4]:4d} smem={r[5]:>6s}  {n}")
