"""Per-kernel SASS census of libembclip_b200.so: counts of the mnemonics that prove (or disprove) a
Blackwell-native kernel -- UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG / UBLKCP (TMA), LDTM / STTM (tcgen05.ld / st), HMMA (legacy
mma.sync: must be 0), FFMA (CUDA-core fp32).  Runs in the dev container (cuobjdump needs no GPU).
  python tools/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "embodied-clip_b200", "libembclip_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "FFMA", "total"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    it = iter(names)
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it, m.group(1))
            cur = re.sub(r"\((?:int|bool|unsigned int)\)", "", cur)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("embclip::", "")
            per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            per[cur]["total"] += 1
            base = op.split(".")[0]
            if base in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "FFMA"):
                per[cur][base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                per[cur]["UTCHMMA.2CTA"] += 1
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), one row per kernel")
    print(f"# {'kernel':78s} " + " ".join(f"{k:>12s}" for k in KEYS))
    tot = collections.Counter()
    for k, c in sorted(per.items()):
        print(f"{k[:80]:80s} " + " ".join(f"{c[x]:12d}" for x in KEYS))
        tot.update(c)
    print(f"{'TOTAL':80s} " + " ".join(f"{tot[x]:12d}" for x in KEYS))
    tc = [k for k, c in per.items() if c["UTCHMMA"]]
    print(f"# {len(per)} kernels; {len(tc)} issue tcgen05.mma; legacy HMMA instructions: {tot['HMMA']}")


if __name__ == "__main__":
    sys.exit(main())
