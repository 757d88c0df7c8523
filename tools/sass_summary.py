"""SASS instruction-count summary per kernel of libfod_b200.so (runs without a GPU):
    python tools/sass_summary.py > profiles/r2_sass_summary.md
Counts the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier, plus the memory and
barrier instructions, registers and static shared memory per kernel (cuobjdump -res-usage)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "faster_orefsdet_b200", "libfod_b200.so")
GROUPS = [("UTC*MMA", r"^UTC\w*MMA"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"),
          ("UTCBAR", r"^UTCBAR"), ("SYNCS", r"^SYNCS"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
          ("ATOM*", r"^(ATOM|RED)"), ("BAR", r"^BAR"), ("UCGABAR", r"^UCGABAR"), ("LDC*", r"^(LDC|ULDC)"), ("SHFL/VOTE", r"^(SHFL|VOTE|MATCH)"),
          ("FFMA*", r"^(FFMA|FMUL|FADD)"), ("HFMA2/F2F", r"^(HFMA2|HADD2|HMUL2|F2F|F2FP)")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    print("# SASS summary of `libfod_b200.so` (sm_100a), round 2\n")
    print("`python tools/sass_summary.py` (cuobjdump -sass / -res-usage, no GPU needed).  `UTC*MMA` = tcgen05.mma, `LDTM` / `STTM` = "
          "tcgen05.ld / st, `UTMALDG` / `UTMASTG` = TMA load / store, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, "
          "`UCGABAR` = cluster barrier.\n")
    head = ["kernel", "instr", "regs", "static smem"] + [g for g, _ in GROUPS]
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
    for k, cnt in kernels.items():
        total = sum(cnt.values())
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("fod::", "")
        row = [f"`{short}`", str(total)] + [str(v) for v in usage.get(k, ("?", "?"))]
        for _, pat in GROUPS:
            row.append(str(sum(v for op, v in cnt.items() if re.match(pat, op))) or "0")
        print("| " + " | ".join(row) + " |")


if __name__ == "__main__":
    main()
