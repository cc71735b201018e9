"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = r["Kernel Name"].split("(")[0].replace("void ", "")
    agg[n][0] += 1
    agg[n][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.1f} ms total (ncu: cold-cache, serialised - compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={v[0]:4d} total {v[1] / 1e6:10.3f} ms avg {v[1] / v[0] / 1e3:10.1f} us share {100 * v[1] / tot:6.2f}%")
