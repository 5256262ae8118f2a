"""Hottest SASS instructions of an `ncu --page source --csv` export (stall samples + executed)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    i_src, i_smp, i_exe = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    body = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break                      # first launch only
        if len(r) > i_exe and r[0] != 'Address':
            body.append(r)
    tot = sum(int(r[i_smp] or 0) for r in body)
    tot_exe = sum(int(r[i_exe] or 0) for r in body)
    print('total samples %d, warp instructions %d' % (tot, tot_exe))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][i_smp] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        print('%5d %5.1f%% exe=%9s  %s' % (i, 100.0 * int(r[i_smp] or 0) / max(tot, 1), r[i_exe], r[i_src].strip()[:110]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
