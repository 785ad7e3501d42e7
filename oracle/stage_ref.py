#!/usr/bin/env python
"""Stage the reference's own Python modules next to the oracle so that they can be TIMED on the GPU box.

    python oracle/stage_ref.py          (run by __graft_entry__.build() in the build container)

The reference is pure Python (no C to compile into a library), and /root/reference does not exist on the GPU
box.  BASELINE.md §4 asks for the reference's `Scanner.scan` and its calibration loop timed on the box's own host
cores beside the GPU numbers, so this recipe copies the files those two loops import — lib/*.py, rmcalibrate.py and
the four shipped models — into oracle/_ref/, which is git-ignored (nothing of the reference enters the history) but
travels to the box with the working tree.  Only oracle/ref_bench.py imports from there; the product never does.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RMD_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage():
    if not os.path.isdir(os.path.join(REF, "lib")):
        print("stage_ref: %s is not here; keeping %s as it is" % (REF, DST))
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "lib"))
    os.makedirs(os.path.join(DST, "models"))
    for f in sorted(os.listdir(os.path.join(REF, "lib"))):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(REF, "lib", f), os.path.join(DST, "lib", f))
    shutil.copyfile(os.path.join(REF, "rmcalibrate.py"), os.path.join(DST, "rmcalibrate.py"))
    for f in sorted(os.listdir(os.path.join(REF, "models"))):
        shutil.copyfile(os.path.join(REF, "models", f), os.path.join(DST, "models", f))
    os.system("chmod -R u+w %s" % DST)
    print("stage_ref: staged %s -> %s" % (REF, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
