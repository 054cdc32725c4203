"""Recipe: place the UNMODIFIED reference hot path under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference is pure Python/PyTorch, so "building" it means copying the files of the path byte for
byte from /root/reference (build container only) into oracle/_ref/, which is git-ignored (no reference
source enters the history) but NOT gpurun-ignored, so it travels to the GPU box like a built `.so`.
bench.py's `--impl reference` / `cpu_baseline` legs then time the reference's own code
(`cpu_baseline.kind: "reference"`); when oracle/_ref is absent they fall back to the oracle port.

Usage:  python oracle/build_ref.py [/root/reference]
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

# files of the hot path plus what their imports need
FILES = [
    "metrics.py",
    "models/__init__.py", "models/utils.py", "models/SSD300.py", "models/SSD512.py", "models/RetinaNet.py",
    "models/RefineDet512.py", "models/FCOSDet.py",
    "operators/__init__.py", "operators/Loss.py", "operators/iou_utils.py", "operators/Deformable_convolution.py",
    "dataset/__init__.py", "dataset/transforms.py",
    "detect_scripts/__init__.py", "detect_scripts/detect_tools.py",
]


def build(ref_root="/root/reference", quiet=False):
    """Copy the files; returns DEST, or None when the reference is not available here."""
    if not os.path.isdir(ref_root):
        return DEST if os.path.isdir(DEST) else None
    for rel in FILES:
        src, dst = os.path.join(ref_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if not quiet:
        print("oracle/_ref: %d reference files in place" % len(FILES))
    return DEST


def available():
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
