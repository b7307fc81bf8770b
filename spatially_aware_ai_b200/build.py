"""Builds spatially_aware_ai_b200/libsaf_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree and explicit (no JIT cache): the built .so travels with the repo snapshot to the GPU box.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only build container.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(HERE, "libsaf_b200.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
                "-I", CSRC]
# saf_fusion.cu reproduces the reference's fp32 roundings: never contract mul+add there.
PER_FILE_FLAGS = {"saf_fusion.cu": ["-fmad=false"], "saf_mesh.cu": ["-fmad=false"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libsaf_b200.so")


def _host_compiler_flags():
    # the image exports CC/CXX pointing at a trimmed gcc; nvcc needs a complete host g++
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/saf_b200.h"]:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and (name.endswith((".cu", ".cuh", ".h"))):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(repr((ARCH_FLAGS, COMMON_FLAGS, PER_FILE_FLAGS)).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: an experimental build of the same ABI next to the product library
    (libsaf_b200_<variant>.so, compiled with -D<define>...; selected at run time with SAF_LIB_PATH for A/B timing)."""
    obj_dir = OBJ_DIR if not variant else os.path.join(OBJ_DIR, variant)
    lib_path = LIB_PATH if not variant else os.path.join(HERE, "libsaf_b200_%s.so" % variant)
    stamp = os.path.join(obj_dir, "fingerprint.txt")
    fp = _fingerprint() + repr(tuple(defines))
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == fp:
        return lib_path
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        cmd = [nvcc] + _host_compiler_flags() + ARCH_FLAGS + COMMON_FLAGS + PER_FILE_FLAGS.get(src, []) + \
              ["-D" + d for d in defines] + \
              (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("building libsaf_b200.so failed")
    cmd = [nvcc] + _host_compiler_flags() + ARCH_FLAGS + ["-shared", "-o", lib_path] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl"]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as f:
        f.write(fp)
    return lib_path


if __name__ == "__main__":
    # python -m spatially_aware_ai_b200.build [--force] [-v] [--variant NAME -DNAME=VALUE ...]
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, defines=defs))
