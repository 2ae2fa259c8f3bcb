"""oracle/make_ref.py -- stage the UNMODIFIED reference package for the GPU box.

    python oracle/make_ref.py

The reference's structured-inference package (/root/reference/src/model/torch_struct) is pure Python and needs only
torch.  /root/reference does not exist on the GPU box, so ``__graft_entry__.build()`` runs this recipe in the build
container: it copies the package's ``*.py`` files, byte for byte, into ``oracle/_ref/torch_struct/``.  That directory
is listed in .gitignore (no reference source enters the repository's history) but is NOT gpurun-ignored, so it travels
to the box exactly as the built ``.so`` files do.  Users: ``bench.py --impl reference`` (the reference arm: the
reference's own PyTorch path timed on the box's host cores) and the live-reference parity tests.
"""
import filecmp
import os
import shutil
import sys

SRC = "/root/reference/src/model/torch_struct"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "torch_struct")


def make(verbose=True):
    """Copy the package if the reference tree is present; return the staged directory or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    n = 0
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            out_dir = os.path.join(DST, rel) if rel != "." else DST
            os.makedirs(out_dir, exist_ok=True)
            src, dst = os.path.join(root, f), os.path.join(out_dir, f)
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
                n += 1
    if verbose and n:
        print(f"[make_ref] staged {n} file(s) of the reference's torch_struct package in {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
