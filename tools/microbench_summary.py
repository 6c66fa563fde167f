"""Render tools/microbench.py JSON lines as the text table kept under profiles/ (fixed cases + the configs[4] sweep summary).

    python tools/microbench_summary.py fixed.jsonl sweep.jsonl > profiles/r2_microbench.txt
"""
import json
import re
import statistics
import sys


def load(path):
    out = []
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            try:
                out.append(json.loads(line))
            except ValueError:
                pass
    return out


def clocks(d):
    c = d.get("clocks", {})
    reasons = ",".join(c.get("reasons", [])) or "-"
    return "[sm %s MHz %s %s]" % (c.get("sm_mhz"), reasons, "graph" if c.get("graph_replay") else "eager")


def main(fixed_path, sweep_path=None):
    fixed = load(fixed_path)
    print("== fixed cases")
    for d in fixed:
        if "ms" not in d:
            print("%-118s %s" % (d.get("case"), d.get("skipped") or d.get("error")))
            continue
        rate = "%8.1f GB/s" % d["achieved_GBs"] if d.get("bound") == "hbm" else "%8.1f TF/s" % d.get("achieved_TFs", 0.0)
        print("%-118s %8.4f ms %s  frac %.3f  %s" % (d["case"], d["ms"], rate, d.get("frac", 0.0), clocks(d)))
    if not sweep_path:
        return
    sweep = [d for d in load(sweep_path) if re.search(r" B=\d+ C=\d+ H=\d+$", d.get("case", ""))]
    done = [d for d in sweep if "ms" in d]
    print()
    print("== configs[4] sweep: C in {64..512} x H in {32..384} x B in {1, 8, 32}, forward and backward (%d lines: %d measured, "
          "%d skipped: tensors beyond --max-gb or the time budget)" % (len(sweep), len(done), len(sweep) - len(done)))
    print("   fraction of the roofline by op and size class (median over the measured shapes; small = under 64 MB of algorithmic "
          "traffic / 20 GFLOP)")
    groups = {}
    for d in done:
        op = re.sub(r" B=\d+ C=\d+ H=\d+$", "", d["case"])
        big = d.get("algorithmic_MB", 0) >= 64 or d.get("algorithmic_GFLOP", 0) >= 20
        groups.setdefault((op, "large" if big else "small"), []).append(d["frac"])
    for (op, cls), v in sorted(groups.items()):
        print("   %-44s %-5s n=%3d  median %.3f  max %.3f" % (op, cls, len(v), statistics.median(v), max(v)))


if __name__ == "__main__":
    main(*sys.argv[1:3])
