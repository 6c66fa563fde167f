"""SASS opcode histogram of libspgan_b200.so (cuobjdump -sass): per kernel, the counts of the instructions that prove the
Blackwell path (UTCHMMA = tcgen05.mma, UTMALDG = TMA loads incl. .IM2COL / .2CTA, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UCGABAR = cluster barriers, FFMA2 / FMUL2 / FADD2 = packed fp32) and the total instruction count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sp-gan-tip2025_b200", "csrc", "libspgan_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMALDG.IM2COL", "UTMALDG.2CTA", "LDTM", "UTCBAR", "UTCBAR.2CTA", "UCGABAR_ARV", "SYNCS", "FFMA2", "FMUL2",
        "FADD2", "HMMA", "LDG.E.128", "STG.E.128", "STS.128", "LDS.128", "UBLKCP", "MUFU"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1).rstrip(";")
            cur["total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if op.startswith("UTMALDG") and "IM2COL" in op:
                cur["UTMALDG.IM2COL"] += 1
            if op.startswith("UTMALDG") and "2CTA" in op:
                cur["UTMALDG.2CTA"] += 1
            if op.startswith("UTCBAR") and "2CTA" in op:
                cur["UTCBAR.2CTA"] += 1
            for k in ("LDG.E.128", "STG.E.128", "STS.128"):
                if op.startswith(k):
                    cur[k] += 1
    tot = collections.Counter()
    print("%-64s %7s  %s" % ("kernel", "instrs", "  ".join(KEYS)))
    for name, c in per.items():
        tot.update(c)
        shown = {k: c[k] for k in KEYS if c[k]}
        print("%-64s %7d  %s" % (name[:64], c["total"], "  ".join("%s=%d" % kv for kv in shown.items())))
    print("\nlibrary totals: " + "  ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot[k]) + "  kernels=%d  instructions=%d" % (len(per), tot["total"]))


if __name__ == "__main__":
    main()
