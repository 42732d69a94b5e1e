"""sum an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name: python tools/exp/launch_sum.py file.csv [skip_first_n]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
t = collections.Counter(); n = collections.Counter()
for r in rows:
	name = re.sub(r'\(.*', '', r[4]); v = float(r[-1]); unit = r[-2]
	if unit in ('us', 'usecond'): v *= 1e3
	elif unit in ('ms', 'msecond'): v *= 1e6
	t[name] += v; n[name] += 1
tot = sum(t.values())
for k, v in t.most_common(25): print("%10.1f us %6d x %8.2f us  %5.1f%%  %s" % (v / 1e3, n[k], v / 1e3 / n[k], 100 * v / tot, k))
print("total %.1f us, %d launches" % (tot / 1e3, len(rows)))
