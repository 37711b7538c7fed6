"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference sources of the hot path.

``/root/reference`` exists only in the build container.  ``stage()`` copies the two source files the oracle
shim executes (``scdrs/pp.py`` and ``scdrs/method.py``; pure Python, nothing to compile) to ``oracle/_ref/scdrs/``.
That directory is git-ignored (the reference's sources never enter this repository's history) but it is not
gpurun-ignored, so -- like the built ``.so`` -- it travels with the working tree to the GPU box, where
``bench.py``'s CPU legs then time the real reference (``cpu_baseline.kind = "reference"``) instead of the numpy
port.  ``__graft_entry__.build()`` runs this when ``/root/reference`` is present.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("scdrs/pp.py", "scdrs/method.py")


def stage():
    """Copy the files when the source tree is there; returns the staged root or None."""
    if not all(os.path.exists(os.path.join(SRC_ROOT, f)) for f in FILES):
        return DST_ROOT if all(os.path.exists(os.path.join(DST_ROOT, f)) for f in FILES) else None
    for f in FILES:
        dst = os.path.join(DST_ROOT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, f), dst)
    return DST_ROOT


if __name__ == "__main__":
    print(stage())
