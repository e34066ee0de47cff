"""ncu launch list -> per-launch table of the LAST face forward (kernel, grid, ms), tc GEMM launches grouped by grid size."""
import csv, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows)); hdr = r[0]
ik, iv, iu, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
L = []
for x in r[1:]:
    v = float(x[iv].replace(",", "")); u = x[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    L.append((x[ik].split("(")[0][:60], x[ig], ms))
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = len(L) // nf
last = L[-n:]
print("launches per forward: %d, total %.2f ms" % (n, sum(m for _, _, m in last)))
agg = collections.OrderedDict()
for k, g, m in last:
    key = (k, g)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += m
for (k, g), (c, m) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-62s grid %-18s x%-3d %8.3f ms  (%.3f each)" % (k, g, c, m, m / c))
