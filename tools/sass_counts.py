#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMA / TMEM paths (cuobjdump -sass of the built library).
    python tools/sass_counts.py > profiles/<round>_sass_tensor_tma_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
                               ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("MUFU.EX2", r"MUFU\.EX2"), ("RED/REDG", r"\bRED(G)?\b"),
                               ("MATCH", r"\bMATCH\b")])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "recommendflow_b200", "librf_b200.so")], capture_output=True,
                          text=True, check=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
        elif cur:
            for k, p in PAT.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    print("SASS mnemonic counts per kernel of recommendflow_b200/librf_b200.so (cuobjdump -sass, sm_100a).")
    print("UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM), UTCBAR = tcgen05.commit,")
    print("SYNCS = mbarrier ops, RED = red.global.add, MATCH = match.any.  Kernels with none of these are omitted.\n")
    for f, c in counts.items():
        if any(c[k] for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "RED/REDG", "MATCH")):
            name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
            if name.startswith("void cub::"):
                name = re.sub(r"<.*", "<...>", name)
            print(f"{name:72s} " + "  ".join(f"{k}={c[k]}" for k in PAT if c[k]))


if __name__ == "__main__":
    main()
