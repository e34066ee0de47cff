"""ncu launch list (gpu__time_duration.sum per launch, --csv) -> markdown table of kernel shares."""
import csv, sys, collections, gzip
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
rows = []
with op(path, "rt") as f:
    for line in f:
        if line.startswith('"'):
            rows.append(line)
r = list(csv.reader(rows))
hdr = r[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for x in r[1:]:
    v = float(x[iv].replace(",", ""))
    u = x[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    name = x[ik]
    short = name.split("(")[0][:70] if not name.startswith("void at::") else "at::" + name.split("at::")[-1][:60]
    tot[short] += ms; cnt[short] += 1
T = sum(tot.values())
print("Total %.1f ms over %d launches.\n" % (T, sum(cnt.values())))
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, v in tot.most_common(16):
    print("| `%s` | %d | %.2f | %.1f %% |" % (k, cnt[k], v, 100 * v / T))
