#!/usr/bin/env python
"""Static summary of dune-multidomain_b200/libmdb200.so (no GPU needed): per kernel the register count, shared memory, local-memory
stack (spills) from `cuobjdump -res-usage`, and the counts of the SASS mnemonics that prove the bulk-async (TMA) / mbarrier / FP64 /
reduction paths (`cuobjdump -sass`, the mnemonics /opt/skills/guides/B200_PROFILING.md lists).  Writes Markdown to stdout:

    python profiles/sass_summary.py > profiles/r02/sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dune-multidomain_b200", "libmdb200.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
CUFILT = "/usr/local/cuda/bin/cu++filt"
MNEMONICS = ("UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "REDG", "ATOMG", "ATOMS", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU")


def demangle(names):
    out = subprocess.run([CUFILT], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    short = []
    for s in out:
        s = re.sub(r"^void ", "", s).replace("(anonymous namespace)", "<unnamed>")
        depth, end = 0, len(s)
        for i, ch in enumerate(s):                      # the name ends at the first "(" outside template brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                end = i
                break
        s = s[:end].replace("(int)", "").replace("(bool)", "").replace("mdb::<unnamed>::", "").replace("mdb::", "")
        short.append(s if len(s) <= 96 else s[:93] + "...")
    return short


def res_usage():
    txt = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows = {}
    name = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            d = dict(re.findall(r"(\w+):(\d+)", line))
            rows[name] = d
            name = None
    return rows


def sass_counts():
    proc = subprocess.Popen([CUOBJDUMP, "-sass", LIB], stdout=subprocess.PIPE, text=True)
    counts = collections.defaultdict(collections.Counter)
    arch = set()
    name = None
    for line in proc.stdout:
        if line.startswith("arch ="):
            arch.add(line.split("=")[1].strip())
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            counts[name]["_total"] += 1
            if op in MNEMONICS:
                counts[name][op] += 1
            if op == "UBLKCP":
                counts[name]["UBLKCP" + m.group(2)] += 1
    proc.wait()
    return counts, arch


def main():
    ru = res_usage()
    sc, arch = sass_counts()
    names = sorted(sc)
    short = dict(zip(names, demangle(names)))
    tot = collections.Counter()
    for n in names:
        tot.update(sc[n])
    print("# Static summary of `libmdb200.so` (cuobjdump, no GPU)\n")
    print("Written by `profiles/sass_summary.py`.  Architectures in the fat binary: %s.  Kernels: **%d**; SASS instructions: %d.\n"
          % (", ".join(sorted(arch)), len(names), tot["_total"]))
    print("Whole library: " + ", ".join("`%s` %d" % (k, tot[k]) for k in MNEMONICS if tot[k]) + ".")
    var = sorted(k for k in tot if k.startswith("UBLKCP."))
    if var:
        print("Bulk-async copy forms: " + ", ".join("`%s` %d" % (k, tot[k]) for k in var) + ".")
    spill = [(short[n], ru[n]) for n in names if n in ru and int(ru[n].get("STACK", 0)) > 0]
    print("\nKernels with a local-memory stack (dynamically indexed local arrays or spills): %d of %d.\n" % (len(spill), len(names)))
    print("| kernel | regs | shared B | stack B | SASS | UBLKCP | SYNCS | DFMA | DADD+DMUL | REDG+ATOMG | LDS | LDG | STG |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n in sorted(names, key=lambda k: short[k]):
        r = ru.get(n, {})
        c = sc[n]
        print("| `%s` | %s | %s | %s | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (
            short[n].replace("|", "\\|"), r.get("REG", "?"), r.get("SHARED", "?"), r.get("STACK", "?"), c["_total"], c["UBLKCP"], c["SYNCS"],
            c["DFMA"], c["DADD"] + c["DMUL"], c["REDG"] + c["ATOMG"], c["LDS"], c["LDG"], c["STG"]))


if __name__ == "__main__":
    sys.exit(main())
