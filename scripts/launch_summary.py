"""Summarise an ncu launch list (gpu__time_duration.sum CSV) by kernel name: count, total us, share."""
import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.Counter(); cnt = collections.Counter()
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
    name = re.sub(r'\(.*', '', re.sub(r'<.*', '', row['Kernel Name']))[:48]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print("%-50s %6s %12s %7s %10s" % ("kernel", "n", "total_us", "share", "us/launch"))
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-50s %6d %12.1f %6.1f%% %10.1f" % (k, cnt[k], v, 100 * v / s, v / cnt[k]))
print("total_us %.1f over %d launches" % (s, sum(cnt.values())))
