"""Stage the reference's UNMODIFIED Python entry points next to the reference build, for the live drop-in test.

TEST INFRASTRUCTURE ONLY.  Copies train.py, render.py, gaussian_renderer/, scene/, utils/, arguments/ from /root/reference
into oracle/_ref/refsrc/ -- git-ignored (never part of the history), but shipped to the GPU box like oracle/_ref/ref_dgr_C.so,
because /root/reference does not exist there.  tests/test_gpu_dropin_live.py runs them, byte for byte as copied, against this
repo's `diff_gaussian_rasterization` (the missing third-party imports are satisfied by tests/ref_stubs/).

Usage:  python oracle/stage_ref.py
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref", "refsrc")
ITEMS = ["train.py", "render.py", "gaussian_renderer", "scene", "utils", "arguments"]


def available() -> bool:
    return os.path.isdir(REF)


def staged() -> bool:
    return os.path.isfile(os.path.join(OUT, "train.py"))


def stage(verbose: bool = True) -> str | None:
    if not available():
        return OUT if staged() else None
    os.makedirs(OUT, exist_ok=True)
    for it in ITEMS:
        src, dst = os.path.join(REF, it), os.path.join(OUT, it)
        if os.path.isdir(src):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copyfile(src, dst)
    # proof of "unmodified": every staged file is byte-identical to its source
    for it in ITEMS:
        src, dst = os.path.join(REF, it), os.path.join(OUT, it)
        if os.path.isdir(src):
            cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
            assert not cmp.diff_files and not cmp.left_only, (it, cmp.diff_files, cmp.left_only)
        else:
            assert filecmp.cmp(src, dst, shallow=False), it
    with open(os.path.join(OUT, "README_STAGED.txt"), "w") as f:
        f.write("Byte-identical copies of the reference's own entry points (train.py, render.py, gaussian_renderer/, scene/, utils/, arguments/),\n"
                "staged by oracle/stage_ref.py from /root/reference for tests/test_gpu_dropin_live.py (the GPU box has no /root/reference).\n"
                "TEST INFRASTRUCTURE: git-ignored (oracle/_ref/), never part of the repository's history, never imported by the product.\n")
    if verbose:
        print("[stage_ref] staged", ", ".join(ITEMS), "->", OUT)
    return OUT


if __name__ == "__main__":
    print(stage() or "reference sources absent and nothing staged")
