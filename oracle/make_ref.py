#!/usr/bin/env python
"""Recipe for oracle/_ref/: a verbatim copy of the reference's Python package, made where the reference lies.

The reference is pure Python (no compiled code), so "building" it means making it importable on the GPU box,
where /root/reference does not exist: this script copies /root/reference/deepsignal/ -> oracle/_ref/deepsignal/
(git-ignored, so the reference's sources never enter this repository's history; not gpurun-ignored, so the
copy travels with the snapshot).  It is run by __graft_entry__.build() in the build container.  The only users
are oracle/ref_host.py (bench.py --impl reference / cpu_baseline) and tests that pin the oracle.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DEEPSIGNAL_REFERENCE", "/root/reference")


def main():
    src = os.path.join(SRC, "deepsignal")
    dst = os.path.join(HERE, "_ref", "deepsignal")
    if not os.path.isdir(src):
        print("make_ref: {} not present, keeping the existing oracle/_ref (if any)".format(src))
        return 0
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(f) for _, _, f in os.walk(dst))
    print("make_ref: copied {} files of the reference package to {}".format(n, dst))
    return 0


if __name__ == "__main__":
    sys.exit(main())
