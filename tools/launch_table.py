"""Print the launch list of an `ncu --metrics gpu__time_duration.sum --csv` log: last N launches or per-kernel totals."""
import csv
import sys
import collections


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
        seq.append((r[ki], ms))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if n:
        for k, ms in seq[-n:]:
            print("%-70s %.3f" % (k[:70], ms))
    else:
        tot = collections.OrderedDict()
        for k, ms in seq:
            a = tot.setdefault(k[:70], [0, 0.0])
            a[0] += 1
            a[1] += ms
        s = sum(v[1] for v in tot.values())
        for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print("%-70s %4d %9.3f %5.1f%%" % (k, c, ms, 100 * ms / s))
        print("total %.3f ms" % s)
