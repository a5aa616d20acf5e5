"""Developer tool: build libcwt_b200 variants whose resident fit kernel is compiled with -DRES_VARIANT=<mask>
(ablations / alternative phase implementations, see csrc/fit_resident.cu) into tools/variants/libcwt_v<mask>.so.
Only fit_resident.cu (or the source named by SRC=iou.cu, ...) is recompiled; the other objects are the product build's. A spec
"mask:-DFOO=1,-DBAR=2" adds compiler flags. Load a variant with CWT_LIB_PATH=..."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from few_shot_seg_cwt_b200 import build as B

def main(masks):
    B.build()
    objdir = os.path.join(B.LIBDIR, "obj")
    src = os.environ.get("SRC", "fit_resident.cu")
    others = [os.path.join(objdir, f) for f in sorted(os.listdir(objdir)) if f.endswith(".o") and f != src[:-3] + ".o"]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "variants")
    os.makedirs(out, exist_ok=True)
    for m in masks:
        extra = []
        if isinstance(m, tuple):
            m, extra = m
        elif ":" in m:                          # "mask:-DFOO=1,-DBAR=2" -> extra compiler flags, library named after the whole spec
            m, fl = m.split(":", 1)
            extra = fl.split(",")
        tag = m + "".join("_" + e.replace("-D", "").replace("=", "") for e in extra)
        obj = os.path.join(out, f"fit_resident_v{tag}.o")
        r = subprocess.run([B.NVCC, *B.FLAGS, f"-DRES_VARIANT={m}", *extra, "-c", os.path.join(B.CSRC, src), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stderr)
        lib = os.path.join(out, f"libcwt_v{tag}.so")
        r = subprocess.run([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj, *others], capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stderr)
        os.remove(obj)
        print(lib)

if __name__ == "__main__":
    main([a for a in sys.argv[1:]])
