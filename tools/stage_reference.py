"""Stage the few reference files the drop-in proof needs next to the repo for ONE gpurun call (git-ignored, never
committed): tests/test_gpu_net.py::test_reference_selection_net_forward_over_b2m imports them unmodified.

    python tools/stage_reference.py          # copies into ./_refstage (needs /root/reference, i.e. the build container)
    python tools/stage_reference.py --clean
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "_refstage")
FILES = ["models/detection_net.py", "models/resnet.py", "models/iou_nms.py", "models/__init__.py", "utils/util.py",
         "utils/__init__.py"]

if "--clean" in sys.argv:
    shutil.rmtree(DST, ignore_errors=True)
    sys.exit(0)
for rel in FILES:
    src = os.path.join("/root/reference", rel)
    dst = os.path.join(DST, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if os.path.exists(src):
        shutil.copyfile(src, dst)
    elif rel.endswith("__init__.py"):
        open(dst, "w").close()
print("staged into", DST)
