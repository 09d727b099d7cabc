"""Recipe for oracle/_ref/: a snapshot of the UNMODIFIED reference's own Python files for the hot path, taken from
/root/reference in the build container so that the GPU box (which has no /root/reference) can time the real reference
on its host cores (bench.py --impl reference, cpu_baseline.kind = "reference") - the analogue of the
`pip install --target baseline/_ref` of an installable reference (this one is a script directory without
setup.py / pyproject, so it cannot be pip-installed).

TEST / BENCH INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (never part of the history) but travels with gpurun
snapshots like the built .so files.  Nothing under dwc_gan_b200/ imports it.

Usage:  python oracle/make_ref.py        (a no-op when /root/reference is absent)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
# what solver.py imports on the G+D step path (SURVEY.md 8c "files a restatement must follow")
FILES = ["solver.py", "gmm.py", "tools.py", "utils.py", "vocab.py", "networks/__init__.py", "networks/networks.py",
         "networks/networks_v2.py", "configs/celeba_faces.yaml"]


def make_ref(verbose=False):
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                open(os.path.join(DST, rel), "a").close()
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("snapshot", rel)
    return True


def import_reference():
    """(Solver class, utils module, config path) of the snapshot, through the two import stubs of SURVEY.md 8c."""
    import types
    if not os.path.exists(os.path.join(DST, "solver.py")):
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
    for name in ("torchfile", "tensorboardX"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.load = lambda *a, **k: None
            m.SummaryWriter = object
            sys.modules[name] = m
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import utils as ref_utils
    from solver import Solver
    return Solver, ref_utils, os.path.join(DST, "configs", "celeba_faces.yaml")


if __name__ == "__main__":
    print("oracle/_ref ready" if make_ref(verbose=True) else "no /root/reference here: nothing to snapshot")
