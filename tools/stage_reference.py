#!/usr/bin/env python3
"""Stage the reference's own driver scripts next to the tests as a git-ignored TEST ASSET (tests/_ref/), so that the GPU
box -- which has no /root/reference -- can run them UNMODIFIED against the CUDA drop-in (SURVEY.md section 7 step-2 gate,
section 8c claim 3).  Nothing under fhe_spear_b200/ or bench.py reads these files; they are not
part of the repository's history (tests/_ref/ is listed in .gitignore, not in .gpurunignore, so the snapshot carries them
like a built .so).  Run by __graft_entry__.build() whenever /root/reference is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["fhe_common.py", "fhe_rwkv_inference.py", "test_fully_enc_bsgs.py", os.path.join("scripts", "bootstrap_generation.py")]


def stage(quiet=False):
    if not os.path.isdir(REF):
        return False
    dst_root = os.path.join(ROOT, "tests", "_ref")
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    if not quiet:
        print(f"staged {len(FILES)} reference scripts under {dst_root}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
