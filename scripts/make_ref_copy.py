#!/usr/bin/env python
"""Copy the files of the UNMODIFIED reference that tests/test_reference_callers_gpu.py drives (its own callers train.forward and
val_lm.visdial_evaluate, and its own model as the on-device comparison) into the git-ignored ``baseline/_ref/reference`` so that
they travel to the GPU box, where /root/reference does not exist.  Sources only under baseline/_ref (never in history).

    python scripts/make_ref_copy.py [--ref /root/reference]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["train.py", "val_lm.py", "val.py", "options.py", "models", "utils", "dataloader", "config"]


def main(ref="/root/reference"):
    dst = os.path.join(ROOT, "baseline", "_ref", "reference")
    if not os.path.isdir(ref):
        print(f"{ref} not present: nothing copied")
        return False
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    for name in WANT:
        src = os.path.join(ref, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(dst, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.lmdb"))
        elif os.path.exists(src):
            shutil.copy2(src, os.path.join(dst, name))
    print("copied", WANT, "->", dst)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    main(ap.parse_args().ref)
