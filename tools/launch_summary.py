"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        if u in ('nsecond', 'ns'):
            v /= 1e3
        elif u in ('msecond', 'ms'):
            v *= 1e3
        seq.append((name, v))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in seq:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("launches %d, total %.2f ms" % (len(seq), tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
        print("%9.2f ms %5.1f%% %6d  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], k[:90]))
