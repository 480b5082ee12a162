"""Build libai_b200.so (the C-ABI library of include/ai_b200.h) in-tree with nvcc for sm_100a.

    python -m adaptive_interpolation_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The product is adaptive_interpolation_b200/_lib/libai_b200.so
(git-ignored; it travels to the GPU box with the working tree).  Replaces the reference's
per-approximation `ispc` + `g++ -shared` build at run_test.py:46-122,190-218: ONE library, built
once, serves every table.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
OBJROOT = os.path.join(ROOT, "build")          # scratch: git-ignored AND gpurun-ignored (objects never travel)
OBJDIR = os.path.join(OBJROOT, "obj")
LIB = os.path.join(LIBDIR, "libai_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "ai_b200.h"))
    return sorted(hs)


def fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for path in sources() + headers():
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_current():
    stamp = os.path.join(LIBDIR, "fingerprint")
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == fingerprint()


def build(force=False, verbose=False, defines=(), out=None, only=None):
    """Compile every CUDA translation unit for sm_100a and link the shared library.
    `defines` / `out` build a tuning variant (tools/tune.py) next to the product library; `only`
    (base names of .cu files) recompiles just those units with the defines and links them with the
    product's objects for the rest (a variant for ONE kernel family in seconds instead of minutes)."""
    variant = out is not None
    lib = os.path.join(LIBDIR, out) if variant else LIB
    if not variant and not force and is_current():
        return LIB
    nvcc = nvcc_path()
    objdir = os.path.join(OBJROOT, "obj_" + out) if variant else OBJDIR
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)         # git-ignored: absent in a fresh checkout
    dflags = ["-D" + d for d in defines]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + dflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    todo = sources()
    reused = []
    if only is not None:
        if not variant:
            raise ValueError("`only` is for variants")
        todo = [src for src in sources() if os.path.basename(src) in only]
        reused = [os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o") for src in sources() if src not in todo]
        missing = [o for o in reused if not os.path.exists(o)]
        if missing:
            raise RuntimeError("build the product library first (missing %s)" % missing[0])
    workers = max(1, min(len(todo), os.cpu_count() or 1))
    with concurrent.futures.ThreadPoolExecutor(workers) as ex:
        results = list(ex.map(compile_one, todo))
    objs = [o for o, _ in results] + reused
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if not variant:
        with open(os.path.join(LIBDIR, "fingerprint"), "w") as fh:
            fh.write(fingerprint())
    return lib


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
