"""Builds dwc_gan_b200/libdwc_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libdwc_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found")
    return nv


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "dwc_b200.h"))
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    flags = list(NVCC_FLAGS)
    if os.environ.get("DWC_EXPERIMENTAL", "0") not in ("", "0"):        # parked kernels, see csrc/experimental/README.md
        srcs += sorted(os.path.join("experimental", f) for f in os.listdir(os.path.join(CSRC, "experimental"))
                       if f.endswith(".cu"))
        flags.append("-DDWC_EXPERIMENTAL")
    nv = _nvcc()

    def compile_one(f):
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, os.path.basename(f)[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            r = subprocess.run([nv] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
            with open(obj + ".log", "w") as lf:
                lf.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (f, r.stdout + r.stderr))
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB, objs):
        r = subprocess.run([nv, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
                           ["-lcudart_static", "-lpthread", "-ldl", "-lrt"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
