"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel: count, mean ns, share of the total."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki]].append(float(r[vi].replace(',', '')))
unit = rows[1][ui]
tot = sum(sum(v) for v in agg.values())
print('# %s ; unit %s ; total %.1f over %d launches' % (sys.argv[1], unit, tot, sum(len(v) for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-70s n=%5d mean=%10.1f share=%.3f' % (k[:70], len(v), sum(v) / len(v), sum(v) / tot))
