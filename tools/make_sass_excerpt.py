#!/usr/bin/env python
"""Writes profiles/<tag>_sass_excerpt.txt from the in-tree libdrone2d.so (needs only cuobjdump, no GPU):
per kernel the instruction count and the counts of the mnemonics that matter, then every tensor-core /
tensor-memory / bulk-copy instruction of the policy kernel.

    python tools/make_sass_excerpt.py [--tag r02]
"""
import argparse
import os
import re
import subprocess
from collections import Counter, OrderedDict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "drone_2d_custom_gym_env_for_reinforcement_learning_b200", "csrc", "libdrone2d.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTCATOMSWS", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H",
        "MUFU.RCP64H", "LDG", "STG", "LDS", "STS", "LDGSTS", "SHFL", "BAR", "ACQBULK"]
TENSOR = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTCATOMSWS")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r02")
    a = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = OrderedDict(), None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            funcs[cur].append((ln.rstrip(), m.group(2)))
    out = ["# SASS evidence (cuobjdump -sass csrc/libdrone2d.so, sm_100a cubin) -- written by tools/make_sass_excerpt.py",
           "# per kernel: instruction count, and the counts of the mnemonics that prove tcgen05 / TMEM / bulk-copy use", ""]
    for name, ins in funcs.items():
        if not name.startswith("_Z") or "d2d_" not in name or "kernel" not in name:
            continue
        c = Counter()
        for _, t in ins:
            p = t.split()
            op = p[1] if p[0].startswith("@") and len(p) > 1 else p[0]
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    c[k] += 1
        out.append(name)
        out.append("    instructions: %d  %s" % (len(ins), "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
    out += ["", "# d2d_policy_kernel: every tensor-core / tensor-memory / bulk-copy instruction", ""]
    for name, ins in funcs.items():
        if "d2d_policy_kernel" in name:
            out += [ln for ln, t in ins if any(k in t for k in TENSOR)]
    path = os.path.join(REPO, "profiles", "%s_sass_excerpt.txt" % a.tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(out), "lines")


if __name__ == "__main__":
    main()
