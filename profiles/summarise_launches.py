"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python profiles/summarise_launches.py <launches.csv> <title>"""
import collections
import csv
import sys


def main(path, title):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) > iv:
            agg.setdefault(r[ik], []).append(float(r[iv]))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {title}")
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
    print("kernel,launches,mean_us,share_of_listed_time")
    for k, v in agg.items():
        print(f"\"{k}\",{len(v)},{sum(v) / len(v) / 1e3:.1f},{sum(v) / tot:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
