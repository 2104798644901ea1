"""ctypes binding of the msha_b200 C-ABI shared library (``include/msha_b200.h``).

The prototypes are parsed from the header, so the header is the single source of truth for the
boundary.  There is no CPU or eager fallback: if the library is missing the import of any op
fails loudly with build instructions.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
import sys

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG_DIR)
HEADER = os.path.join(_ROOT, "include", "msha_b200.h")
CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libmsha_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"]

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "size_t": ctypes.c_size_t,
}


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` for sm_100a into ``libmsha_b200.so`` (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = list(NVCC_FLAGS) + os.environ.get("MSHA_NVCC_EXTRA", "").split()     # e.g. -DMSHA_DZ_PROD_WARPS=8 (tuning)
    objs = []
    obj_dir = os.path.join(_PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [HEADER]
    stamp = os.path.join(obj_dir, "flags.txt")
    same_flags = os.path.isfile(stamp) and open(stamp).read() == " ".join(flags)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and same_flags and os.path.isfile(obj)
                and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in [src] + headers)):
            continue                                   # object is newer than its source and every header
        cmd = [nvcc, *flags, "-I", os.path.join(_ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out.decode(errors='replace')}")
    with open(stamp, "w") as f:
        f.write(" ".join(flags))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout.decode(errors='replace')}")
    return LIB_PATH


def parse_header(path: str = HEADER):
    """-> list of (name, restype_str, [(ctype_str, argname)]) for every prototype in the header."""
    with open(path) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(msha_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        arglist = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos.append((name, ret, arglist))
    return protos


def _to_ctype(t: str):
    t = t.replace("const", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t.replace(" ", "") == "char*" else ctypes.c_void_p
    return _CTYPES[t]


_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the library with prototypes installed."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"msha_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    for name, ret, args in parse_header():
        fn = getattr(L, name)   # AttributeError here == header/library mismatch: fail loudly
        fn.restype = _to_ctype(ret)
        fn.argtypes = [_to_ctype(t) for t, _ in args]
    _lib = L
    return L


def last_error() -> str:
    return lib().msha_last_error().decode(errors="replace")


def check(rc: int, what: str):
    if rc != 0:
        kind = "invalid argument" if rc < 0 else f"CUDA error {rc}"
        raise RuntimeError(f"msha_b200.{what}: {kind}: {last_error()}")
