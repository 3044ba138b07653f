"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches / total / share / mean."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki][:100], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (n, a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / 1e3 / a[0]))
print("| total | | %.1f | | |" % (tot / 1e3))
