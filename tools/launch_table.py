#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST prove.

usage: launch_table.py launches.csv [--all] [--from-kernel NAME]
The last prove is taken as everything after the last `k_sap_public_rows` launch... (first kernel of phase 1).
"""
import csv, re, sys, collections

def rows(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    for r in csv.DictReader(lines[start:]):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('unnamed>::', '').replace('pm::<', '')
            yield int(r['ID']), name.strip(), r['Grid Size'], float(r['Metric Value'].replace(',', '')) / (1e3 if r['Metric Unit'] in ('ns', 'nsecond') else 1)

def main():
    path = sys.argv[1]
    rs = list(rows(path))
    marks = [i for i, r in enumerate(rs) if 'k_ra_square' in r[1] or 'k_sap_public' in r[1]]
    # proves start at k_ra_square (phase 1); take the last complete one: up to the next mark or the first non-prove kernel
    if '--all' not in sys.argv and marks:
        first = marks[-1]
        # the e2e leg may follow; use the one before last if last is the final prove of the file
        seg = rs[first:]
        # stop at benches that follow the proves
        stop = next((i for i, r in enumerate(seg) if 'k_imad_peak' in r[1] or 'k_fill_fr' in r[1]), len(seg))
        seg = seg[:stop]
    else:
        seg = rs
    tot = collections.OrderedDict()
    for _, name, grid, us in seg:
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += us
    total = sum(v[1] for v in tot.values())
    print("launches %d, total %.3f ms" % (len(seg), total / 1e3))
    for name, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %4d %10.3f ms %5.1f%%" % (name[:40], cnt, us / 1e3, 100 * us / total))
    if '--list' in sys.argv:
        for r in seg:
            print(r)

if __name__ == '__main__':
    main()
