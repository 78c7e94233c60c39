"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py file.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    idx = {k: j for j, k in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        n = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("op3::<unnamed>::", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %d launches, %.1f us of kernel time (per-launch times are cold-cache and serialised)" % (sum(v[0] for v in agg.values()), tot))
    print("# %10s %6s %9s %6s  kernel" % ("total us", "count", "avg us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%12.0f %6d %9.1f %5.1f%%  %s" % (t, c, t / c, 100 * t / tot, n[:110]))


if __name__ == "__main__":
    main()
