#!/usr/bin/env python
"""profiles/sass_summary.py -- per-kernel SASS evidence of the shipped library (cuobjdump -sass / -res-usage):
which kernels carry tcgen05 (UTCHMMA / LDTM), TMA tensor loads (UTMALDG, multicast forms), bulk async copies (UBLKCP),
3-input max (FMNMX3), mixed-precision FMA (FHFMA.BF16), and that no legacy warp-level MMA (HMMA / IMMA / DMMA / *GMMA)
is present anywhere.  Usage: python profiles/sass_summary.py > profiles/r02/sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cadence_rag_b200", "libcadence_dense.so")
MNEMONICS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "FMNMX3", "FHFMA", "SYNCS", "STAS", "HMMA", "IMMA", "DMMA", "HGMMA", "QGMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    kernels, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = {k: 0 for k in MNEMONICS}
            kernels[name]["_instr"] = 0
            kernels[name]["_forms"] = set()
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[name]["_instr"] += 1
        base = op.split(".")[0]
        if base in kernels[name]:
            kernels[name][base] += 1
        if base in ("UTMALDG", "UTCHMMA", "UBLKCP", "LDTM", "FHFMA"):
            kernels[name]["_forms"].add(op)
    pretty = demangle(list(kernels))
    elfs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.split()
    print("library:", os.path.relpath(LIB, ROOT))
    print("cubins :", " ".join(e for e in elfs if "cubin" in e))
    print(f"{'kernel':92s} {'instr':>6s} {'regs':>4s} " + " ".join(f"{k:>7s}" for k in MNEMONICS))
    legacy = 0
    for k in sorted(kernels, key=lambda n: pretty[n]):
        v = kernels[k]
        short = re.sub(r"\(anonymous namespace\)::", "", pretty[k])
        short = re.sub(r"\(.*", "", short)[:92]
        print(f"{short:92s} {v['_instr']:6d} {regs.get(k, 0):4d} " + " ".join(f"{v[m]:7d}" for m in MNEMONICS))
        if v["_forms"]:
            print(" " * 10 + "forms: " + " ".join(sorted(v["_forms"])))
        legacy += sum(v[m] for m in ("HMMA", "IMMA", "DMMA", "HGMMA", "QGMMA"))
    print(f"\nlegacy warp-level MMA instructions in the whole library: {legacy}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
