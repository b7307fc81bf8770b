#!/usr/bin/env python
"""Install the reference's hot-path modules for the CPU arm of bench.py (`--impl reference`).

The reference is pure Python with no build system, so "installing" it is placing its three
hot-path modules - UNMODIFIED - where the GPU box can import them: baseline/_ref/ (git-ignored, shipped by gpurun).
`pip install /root/reference` does not apply: the checkout has no setup.py / pyproject.toml.  Run in the build
container (where /root/reference exists); __graft_entry__.build() calls this.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("clipfusion.py", "clip_seem_fusion.py", "handy_utils.py")


def install(src=SRC, dst=DST):
    if not os.path.isdir(src):
        return None
    os.makedirs(dst, exist_ok=True)
    lines = []
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        with open(os.path.join(dst, name), "rb") as f:
            lines.append("%s  %s" % (hashlib.sha256(f.read()).hexdigest(), name))
    with open(os.path.join(dst, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return dst


if __name__ == "__main__":
    out = install()
    print(out or "no reference checkout at %s; baseline/_ref left as it is" % SRC)
    sys.exit(0)
