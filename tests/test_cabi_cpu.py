"""The C-ABI boundary without a GPU: libdwc_b200.so loads, exports every function include/dwc_b200.h declares, every name
the ctypes binding uses exists, and the ctypes mirrors of the parameter structs have the size and field offsets a C
compiler gives the header's structs (no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dwc_b200.h")
LIB = os.path.join(ROOT, "dwc_gan_b200", "libdwc_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"#ifdef DWC_EXPERIMENTAL.*?#endif", "", src, flags=re.S)      # parked kernels: not in the default build
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dwc_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_function():
    assert os.path.exists(LIB), "build the library first: python -m dwc_gan_b200.build"
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.dwc_last_error.restype = ctypes.c_char_p
    assert lib.dwc_last_error() is not None


def test_binding_uses_only_exported_names():
    from dwc_gan_b200 import _lib as L
    lib = ctypes.CDLL(LIB)
    src = open(os.path.join(ROOT, "dwc_gan_b200", "_lib.py")).read()
    used = set(re.findall(r'"(dwc_[a-z0-9_]+)"', src))
    used = {n for n in used if "cluster" not in n and "halo_ok" not in n}             # optional experimental entry points
    assert len(used) >= 50
    missing = [n for n in sorted(used) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(used) <= set(_declared()) | {"dwc_last_error"}, sorted(set(used) - set(_declared()))
    assert L.GConv and L.WGrad and L.HBuf and L.PackEntry and L.AdvTerm


def test_struct_layouts_match_the_header():
    from dwc_gan_b200 import _lib as L
    pairs = [("dwc_gconv_t", L.GConv), ("dwc_wgrad_t", L.WGrad), ("dwc_hbuf_t", L.HBuf),
             ("dwc_pack_entry_t", L.PackEntry), ("dwc_adv_term_t", L.AdvTerm)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void) {"]
    for cname, ct in pairs:
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for f in ct._fields_:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, f[0]))
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "abi.c"), os.path.join(td, "abi")
        open(src, "w").write("\n".join(lines))
        r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe],
                           capture_output=True, text=True)
        if r.returncode != 0 and "cuda" in r.stderr.lower():
            pytest.skip("header needs CUDA types to compile with plain gcc: " + r.stderr[:200])
        assert r.returncode == 0, r.stderr
        out = subprocess.run([exe], capture_output=True, text=True).stdout.strip().splitlines()
    for (cname, ct), line in zip(pairs, out):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == ctypes.sizeof(ct), (cname, parts[1], ctypes.sizeof(ct))
        for f, off in zip(ct._fields_, parts[2:]):
            assert int(off) == getattr(ct, f[0]).offset, (cname, f[0], off, getattr(ct, f[0]).offset)
