"""Instruction census of the built library: per kernel, how many tcgen05 / TMEM / TMA instructions its SASS holds
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk, HMMA
= legacy mma.sync).  CPU only (cuobjdump):   python tools/sass_census.py > profiles/rNN_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dwc_gan_b200", "libdwc_b200.so")
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                               ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
                               ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("FFMA", r"\bFFMA")])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, pat in PAT.items():
            if re.search(pat, line):
                kernels[cur][k] += 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, dem)) if len(dem) == len(kernels) else {k: k for k in kernels}
    print("# SASS instruction census of dwc_gan_b200/libdwc_b200.so (%s; cuobjdump -sass), %d kernels" % (
        ", ".join(sorted(arch)), len(kernels)))
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---|" * len(PAT))
    rows = sorted(kernels.items(), key=lambda kv: (-kv[1]["UTC*MMA"], -kv[1]["UTMALDG"] - kv[1]["UBLKCP"], names[kv[0]]))
    for k, c in rows:
        nm = names[k].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        nm = re.sub(r"\((int|bool|unsigned int)\)", "", nm)
        nm = re.sub(r"\(.*", "", nm)
        print("| `%s` | " % nm[:90] + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + " |")
    tc = [k for k, c in kernels.items() if c["UTC*MMA"]]
    tma = [k for k, c in kernels.items() if c["UTMALDG"] or c["UBLKCP"]]
    print("\n%d kernels issue tcgen05.mma (UTC*MMA) and read their accumulators from TMEM (LDTM); %d kernels stage data "
          "with TMA (UTMALDG tensor copies / UBLKCP bulk copies); no kernel uses legacy HMMA: %s." % (
              len(tc), len(tma), "true" if not any(c["HMMA"] for c in kernels.values()) else "FALSE"))


if __name__ == "__main__":
    main()
