"""Builds libabm_b200.so in-tree with nvcc for sm_100a (the only target).  No torch involvement:
the library is a plain C-ABI shared object loaded with ctypes."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libabm_b200.so")
SOURCES = ["lib.cu", "sort.cu", "search.cu", "jkr.cu", "jkr32.cu", "diffusion.cu", "diffusion3d.cu", "slab.cu", "rules.cu"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-fvisibility=hidden"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libabm_b200.so cannot be built")


STAMP = LIB + ".sha256"          # hash of the sources + flags the in-tree library was built from


def source_hash(defines=()):
    """sha256 over every file of csrc/, the public header, this recipe's flags and the -D switches."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + \
        [os.path.join(ROOT, "include", "abm_b200.h"), os.path.join(ROOT, "include", "abm_rng.h")]
    for d in deps:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(repr((FLAGS, sorted(defines))).encode())
    return h.hexdigest()


def needs_build():
    """True unless the in-tree library was built from exactly the sources that are in the tree now (a hash, not a
    modification time: a checkout, a copy to another box or a touched file must not fool it)."""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build a variant library with -D switches (kernel A/B experiments, scripts/ab_substep.py)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-shared", "-o", out or LIB] + FLAGS + [
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-Xptxas", "-v" if verbose else "-warn-spills"] + [os.path.join(CSRC, s) for s in SOURCES]
    # the image's CC wrapper lacks pieces nvcc's host pass wants; use the distro g++ explicitly
    cmd += ["-D" + d for d in defines]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    if out is None:
        with open(STAMP, "w") as f:
            f.write(source_hash() + "\n")
    return out or LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
