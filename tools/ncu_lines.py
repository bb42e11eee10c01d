#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: executed warp-instructions
and stall samples per CUDA source line of one kernel.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep _Z15d2d_step_kernelILb1EEv9D2dParams [top]

Needs: ncu, cuobjdump, nvdisasm (all in the CUDA toolkit; no GPU).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "drone_2d_custom_gym_env_for_reinforcement_learning_b200", "csrc", "libdrone2d.so")


def disasm_lines(mangled):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    out, cur, active = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = (".text." + mangled + " ") in ln + " "
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def ncu_sass(rep, kernel_substr):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    launches, cur, hdr = [], None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            launches.append(cur)
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if cur is not None and r[0].startswith("0x"):
            cur["rows"].append(r)
    return hdr, [l for l in launches if kernel_substr in l["name"]]


def main():
    rep, mangled = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    dis = disasm_lines(mangled)
    kname = "d2d_step_kernel" if "step" in mangled else ("d2d_reset_kernel" if "reset" in mangled else "")
    hdr, launches = ncu_sass(rep, kname)
    want_true = "Lb1" in mangled
    launches = [l for l in launches if ("(bool)1" in l["name"]) == want_true] or launches
    L = launches[0]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ite = hdr.index("Thread Instructions Executed")
    base = int(L["rows"][0][ia], 16)
    per_line = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in L["rows"]:
        off = int(r[ia], 16) - base
        line, sass = dis.get(off, (None, "?"))
        ex, sm, te = int(float(r[ie])), int(float(r[isamp])), int(float(r[ite]))
        p = per_line[line]
        p[0] += ex; p[1] += sm; p[2] += 1; p[3] += te
        tot[0] += ex; tot[1] += sm; tot[2] += te
    print("kernel %s: %d SASS instrs, %d warp-instr executed, %d stall samples, avg active threads %.1f" %
          (L["name"], len(L["rows"]), tot[0], tot[1], tot[2] / max(tot[0], 1)))
    print("%-28s %8s %6s %6s %6s %5s" % ("line", "warp-ins", "%ins", "%smp", "sass", "thr"))
    for line, (ex, sm, cnt, te) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-28s %8d %6.2f %6.2f %6d %5.1f" % ("%s:%s" % line if line else "?", ex, 100.0 * ex / tot[0],
                                                   100.0 * sm / max(tot[1], 1), cnt, te / max(ex, 1)))


if __name__ == "__main__":
    main()
