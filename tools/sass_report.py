#!/usr/bin/env python
"""Static evidence for the built library: per kernel, registers / spills / shared memory (cuobjdump -res-usage) and the
count of the SASS mnemonics that identify the Blackwell paths (B200_PROFILING.md "What proves a Blackwell-native kernel":
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load / store / reduce,
HMMA = mma.sync, LDSM = ldmatrix).  No GPU needed.
    python tools/sass_report.py [lib] > profiles/<tag>_sass_resources.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gif-gan_b200", "lib", "libgifgan.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDSM", "SYNCS", "ELECT", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("gg::", "")
        n = re.sub(r"\(.*", "", n)
        short.append(n if len(n) < 64 else n[:61] + "...")
    return dict(zip(names, short))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            f = dict(kv.split(":") for kv in line.split() if ":" in kv)
            usage[cur] = f
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                for k in MNEMONICS:
                    if op.startswith(k):
                        counts[cur][k] += 1
                counts[cur]["_n"] += 1
    names = demangle(sorted(usage))
    print("# Static resource usage and Blackwell SASS mnemonics of `%s`\n" % os.path.relpath(LIB, ROOT))
    print("From `cuobjdump -res-usage` and `cuobjdump -sass` (sm_100a).  STACK / LDL / STL > 0 would mean local-memory spills;")
    print("UTCHMMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add,")
    print("HMMA = `mma.sync`, LDSM = `ldmatrix`, SYNCS = mbarrier ops.\n")
    print("| kernel | regs | stack | static smem | SASS instr | " + " | ".join(MNEMONICS) + " |")
    print("|---|---:|---:|---:|---:|" + "---:|" * len(MNEMONICS))
    def key(n):
        c = counts[n]
        return (-(c["UTCHMMA"] > 0), -(c["HMMA"] > 0), names[n])
    shown = [n for n in usage if counts[n]["UTCHMMA"] or counts[n]["HMMA"] or counts[n]["UTMALDG"] or usage[n].get("STACK", "0") != "0"
             or any(t in names[n] for t in ("colsum", "bn_train_apply", "bn_bwd_apply", "adam", "pack_", "distance_loss", "thin_", "lstm_step"))]
    for n in sorted(shown, key=key):
        u, c = usage[n], counts[n]
        print("| `%s` | %s | %s | %s | %d | " % (names[n], u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), c["_n"])
              + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS) + " |")
    rest = [n for n in usage if n not in shown]
    print("\n(+ %d further instantiations -- SIMT parity-mode convs, activations, casts, skinny linears -- none with a stack frame; max %d registers)" % (
        len(rest), max([int(usage[n].get("REG", 0)) for n in rest] or [0])))
    tc = [n for n in usage if counts[n]["UTCHMMA"]]
    print("\n%d kernels; %d issue tcgen05.mma, %d use TMA tensor copies, %d use mma.sync; kernels with a stack frame: %s" % (
        len(usage), len(tc), sum(1 for n in usage if counts[n]["UTMALDG"] or counts[n]["UTMASTG"]),
        sum(1 for n in usage if counts[n]["HMMA"]), ", ".join("`%s`" % names[n] for n in usage if usage[n].get("STACK", "0") != "0") or "none"))


if __name__ == "__main__":
    main()
