"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [--seq N]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    seq = []
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("tbd::<unnamed>::", "").replace("unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        seq.append((k, ms))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k, ms in seq:
        t = tot[k]
        t[0] += 1; t[1] += ms; t[2] = max(t[2], ms)
    T = sum(v[1] for v in tot.values())
    print("| kernel | launches | total ms | share | max ms |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.3f | %.1f %% | %.3f |" % (k, v[0], v[1], 100 * v[1] / T, v[2]))
    if "--seq" in sys.argv:
        n = int(sys.argv[sys.argv.index("--seq") + 1])
        for k, ms in seq[:n]:
            print("   %-28s %.3f" % (k, ms))
