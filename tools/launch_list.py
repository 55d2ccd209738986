"""Summarise an ncu gpu__time_duration launch list: python tools/launch_list.py file.csv [last_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {n: i for i, n in enumerate(hdr)}
seq = [(r[ix['Kernel Name']], float(r[ix['Metric Value']].replace(',', '')), r[ix['Metric Unit']]) for r in rows[1:]]
def us(v, unit):
    return v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
last = int(sys.argv[2]) if len(sys.argv) > 2 else len(seq)
tot = collections.OrderedDict()
for n, v, u in seq[-last:]:
    k = n[:90]
    tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += us(v, u)
all_us = sum(v[1] for v in tot.values())
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%9.1f us %5.1f%% x%-4d %s" % (t, 100 * t / all_us, c, k))
print("total %.1f us over %d launches" % (all_us, sum(v[0] for v in tot.values())))
