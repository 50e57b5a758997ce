# summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3}.get(unit, 1e-6)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print('# %s: %d launches, %.2f ms total device time (cold-cache, serialised under ncu: compare SHARES)' % (sys.argv[1], sum(cnt.values()), T))
print('%-48s %9s %12s %7s' % ('kernel', 'launches', 'time [ms]', 'share'))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print('%-48s %9d %12.3f %6.1f%%' % (k[:48], cnt[k], v, 100 * v / T))
