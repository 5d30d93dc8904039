"""TEST / BENCH INFRASTRUCTURE ONLY - stage the reference's own hot-path modules under oracle/_ref/ so that they can
travel to the GPU box (which has no /root/reference) and be timed there as the CPU baseline.

    python oracle/build_ref.py          # needs /root/reference; writes oracle/_ref/faster_rcnn/{rpn,utils,config,...}.py

The reference is pure Python: "building" it is copying the unmodified files (the same three modules
`oracle/reference_import.py` imports in place in the build container; `augmentation.py` is imported by utils.py).
oracle/_ref/ is git-ignored - no reference source enters the history - but not gpurun-ignored.
`oracle/reference_import.py` falls back to this copy when /root/reference is absent.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("RADNET_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["faster_rcnn/__init__.py", "faster_rcnn/rpn.py", "faster_rcnn/utils.py", "faster_rcnn/config.py",
         "faster_rcnn/augmentation.py"]


def build():
    if not os.path.isfile(os.path.join(SRC, "faster_rcnn", "rpn.py")):
        return None
    for rel in FILES:
        src = os.path.join(SRC, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
    return DST


if __name__ == "__main__":
    out = build()
    print(out if out else "reference tree not found at %s" % SRC)
    sys.exit(0 if out else 1)
