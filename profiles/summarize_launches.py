"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print('total %.1f us over %d launches' % (total, sum(cnt.values())))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%-72s n=%6d tot=%10.1f us avg=%8.2f us share=%5.1f%%'
              % (k[:72], cnt[k], v, v / cnt[k], 100 * v / total))


if __name__ == '__main__':
    main(sys.argv[1])
