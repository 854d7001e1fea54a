"""Hottest SASS instructions of one kernel in an ncu report, by warp-stall samples.

    python tools/ncu_hot.py report.ncu-rep k_train_fwd_tc [top=40]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    if not starts:
        print('kernel not found')
        return
    seg = rows[starts[0] + 1:(starts[1] if len(starts) > 1 else len(rows))]
    hdr, data = seg[0], seg[1:]
    ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[isamp] or 0) for r in data)
    print(rows[starts[0]][1], 'samples', tot, 'instructions', len(data))
    agg = {}
    for r in data:
        for c in stall_cols:
            agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c] or 0)
    print(sorted(agg.items(), key=lambda x: -x[1])[:8])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top_n]
    for i in sorted(top):
        r = data[i]
        st = sorted([(hdr[c], int(r[c] or 0)) for c in stall_cols], key=lambda x: -x[1])[:2]
        print('%5d %5s %7s  %-72s %s' % (i, r[isamp], r[iex], r[ia].strip()[:72], st))


if __name__ == '__main__':
    main()
