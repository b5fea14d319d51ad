"""Builds oracle/_ref/salmetric from the reference's OWN source (compiled where it lies under /root/reference, nothing copied):
/root/reference/CSNet_training/SalMetric/src/sal_metric.cpp + include/sal_metric.hpp, with oracle/cvshim standing in for the three
OpenCV headers (the build container has no OpenCV; the reference's cmake tree is not run).  Test infrastructure only: the binary
pins oracle/salmetric.py (tests/test_salmetric.py).  No-op (returns None) when /root/reference is absent — the GPU box uses the
prebuilt binary that travels with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/CSNet_training/SalMetric"
OUT = os.path.join(HERE, "_ref", "salmetric")


def build(force: bool = False):
    if not os.path.isdir(SRC):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force:
        return OUT
    gxx = shutil.which("g++")
    if gxx is None:
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [gxx, "-O2", "-std=c++11", "-w", "-I", os.path.join(HERE, "cvshim"), "-I", os.path.join(SRC, "include"),
           os.path.join(SRC, "src", "sal_metric.cpp"), "-lpthread", "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building oracle/_ref/salmetric failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
