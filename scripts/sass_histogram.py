"""SASS evidence for profiles/: per kernel of libreacher_b200.so, the instruction count and a histogram of the mnemonics that matter on sm_100a
(UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = 1-D bulk TMA, UTMALDG / UTMASTG = tensor TMA, LDGSTS = cp.async,
FFMA2 / FADD2 / FMUL2 = packed fp32, MUFU.*, BAR / SYNCS = barriers / mbarriers).  Runs here (no GPU): cuobjdump -sass on the built library.
usage: python scripts/sass_histogram.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "reacherdistilation_b200", "libreacher_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "BAR", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL",
        "MUFU.EX2", "MUFU.RCP", "MUFU.TANH", "MUFU.SQRT", "MUFU.RSQ", "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "RED", "MEMBAR", "ERRBAR", "NANOSLEEP"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith("MUFU."):
                kernels[cur][".".join(op.split(".")[:2])] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS instruction histogram of `libreacher_b200.so` (`cuobjdump -sass`, sm_100a; static counts per kernel)", "",
           "Mnemonics: UTCHMMA = `tcgen05.mma kind::f16`, LDTM/STTM = `tcgen05.ld/st`, UBLKCP = `cp.async.bulk` (1-D TMA), UTMALDG/UTMASTG = tensor TMA, "
           "LDGSTS = `cp.async`, FFMA2/FADD2/FMUL2 = packed fp32 pairs, SYNCS = mbarrier ops.", "",
           "| kernel | SASS instr | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for (name, c), d in zip(kernels.items(), dem):
        short = re.sub(r"\(.*", "", d).replace("void ", "").replace("rb::", "")
        out.append("| `%s` | %d | " % (short, c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    out += ["", "Library totals: " + ", ".join("%s %d" % (k, tot[k]) for k in KEYS if tot[k])]
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
