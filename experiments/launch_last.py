"""Per-launch durations of the LAST subject in an ncu `--metrics gpu__time_duration.sum --csv` log (cold-cache, serialised)."""
import csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    val = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit == 'ns': val /= 1e3
    elif unit == 'ms': val *= 1e3
    name = re.sub(r'^void ', '', row['Kernel Name']).replace('ukbb::', '')
    name = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', name)
    seq.append((name, val))
firsts = [i for i, (k, v) in enumerate(seq) if k.startswith('conv_first')]
heads = [i for i, (k, v) in enumerate(seq) if k.startswith('head_ts')]
a, b = firsts[-1], heads[-1]
pre0 = heads[-2] + 1 if len(heads) > 1 else 0
tot_pre = sum(v for k, v in seq[pre0:a]); tot = sum(v for k, v in seq[a:b + 1])
print('--- last subject: preprocessing %d launches %.1f us, forward %d launches %.1f us' % (a - pre0, tot_pre, b + 1 - a, tot))
for k, v in seq[a:b + 1]:
    print('%-90s %9.1f us %5.1f%%' % (k[:90], v, 100 * v / tot))
