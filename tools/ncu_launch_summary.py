"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = r[ki].split("(")[0]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print("%-44s %8s %12s %7s %10s" % ("kernel", "launches", "total ms", "share", "avg us"))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-44s %8d %12.3f %6.1f%% %10.1f" % (n[:44], c, t / 1e6, 100 * t / tot, t / c / 1e3))
print("%-44s %8d %12.3f" % ("total", sum(v[0] for v in agg.values()), tot / 1e6))
