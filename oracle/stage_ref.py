"""TEST INFRASTRUCTURE — stages the reference's own Python sources for this path into oracle/_ref/ (git-ignored; it travels
to the GPU box with the snapshot like a built .so) so that `bench.py --impl reference` and the CPU-baseline leg can time the
UNMODIFIED reference on the box's host cores (`cpu_baseline.kind = "reference"`).  Nothing is copied into tracked files:
oracle/_ref/ is an output directory, rebuilt by `__graft_entry__.build()` whenever /root/reference is mounted.

    python -m oracle.stage_ref
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/icepy4d"
DST = os.path.join(ROOT, "oracle", "_ref", "src", "icepy4d")


def stage(force: bool = False) -> str | None:
    """Mirror the reference package's .py files (1.1 MB) below oracle/_ref/src/.  Returns the staged root or None when the
    reference tree is not mounted (the GPU box: the previously staged copy, if any, is used as is)."""
    if not os.path.isdir(SRC):
        return os.path.dirname(DST) if os.path.isdir(DST) else None
    stamp = os.path.join(os.path.dirname(DST), ".staged")
    newest = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(SRC) for f in fs if f.endswith(".py"))
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= newest:
        return os.path.dirname(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for d, _, fs in os.walk(SRC):
        rel = os.path.relpath(d, SRC)
        for f in fs:
            if f.endswith(".py"):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), os.path.join(DST, rel, f))
    with open(stamp, "w") as fh:
        fh.write("staged from /root/reference/src/icepy4d by oracle/stage_ref.py\n")
    return os.path.dirname(DST)


if __name__ == "__main__":
    print(stage(force=True))
