"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table."""
import collections
import csv
import sys


def main(path, out, title, cmd):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nCommand: `%s`\n(cold-cache, serialised: compare shares, not absolutes)\n\n" % (title, cmd))
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.2f%% |\n" % (k, n, v / 1e6, 100 * v / tot))
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:5])
