"""Recipe for oracle/_ref: the reference's OWN modules for this path, taken verbatim from the checkout where it lies.

    python oracle/make_ref.py            # needs /root/reference (this container); outputs only into oracle/_ref/

The reference is pure Python: "building" it means making its three source files importable without its package
``__init__`` (which pulls in the fork's unrelated models and their dependencies):
    src/model/transformer.py  (MultiHeadAttentionOne, ScaledDotProductAttention)   -> oracle/_ref/ref_src/model/transformer.py
    src/model/conv4d.py       (imported by transformer.py, unused on the path)     -> oracle/_ref/ref_src/model/conv4d.py
    src/util.py               (batch_intersectionAndUnionGPU, intersectionAndUnionGPU) -> oracle/_ref/ref_src/util.py
oracle/_ref/ is git-ignored (reference sources never enter the history) but travels to the GPU box with the snapshot, so
``bench.py --impl reference`` and the CPU baseline there run the literal reference modules (`cpu_baseline.kind` = "reference").
TEST INFRASTRUCTURE ONLY: nothing under few_shot_seg_cwt_b200/ may import it."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CWT_REFERENCE", "/root/reference")
FILES = {"src/model/transformer.py": "ref_src/model/transformer.py", "src/model/conv4d.py": "ref_src/model/conv4d.py",
         "src/util.py": "ref_src/util.py"}


def make(dst_root: str = os.path.join(HERE, "_ref")) -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        return False
    manifest = {}
    for src, dst in FILES.items():
        d = os.path.join(dst_root, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), d)
        manifest[src] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    for pkg in ("ref_src", "ref_src/model"):
        open(os.path.join(dst_root, pkg, "__init__.py"), "w").close()        # empty: do NOT run the reference's package __init__
    json.dump({"reference": REF, "sha256": manifest}, open(os.path.join(dst_root, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref written" if ok else f"{REF} not present: nothing written")
    sys.exit(0 if ok else 1)
