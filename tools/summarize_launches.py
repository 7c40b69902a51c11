"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/r01_launches_summary.txt
"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt, per = collections.defaultdict(float), collections.Counter(), []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = row["Kernel Name"].split("(")[0]
        tot[name] += v
        cnt[name] += 1
        per.append((int(row["ID"]), name, row["Grid Size"], row["Block Size"], v))
    total = sum(tot.values())
    print(f"# {len(per)} launches, {total:.1f} us total device time (cold-cache, serialised by ncu: compare SHARES)")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k:32s} n={cnt[k]:5d} total={tot[k]:10.1f} us  avg={tot[k]/cnt[k]:8.2f} us  share={tot[k]/total:6.1%}")
    print("# id, kernel, grid, block, us")
    for p in per:
        print(f"{p[0]:5d} {p[1]:24s} {p[2]:16s} {p[3]:14s} {p[4]:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
