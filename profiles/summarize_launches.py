"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (B200_PROFILING.md recipe) per kernel:
count, total, share of the step, average.  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except Exception:
            continue
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0].replace('<unnamed>::', '')
        a = agg.setdefault(k, [0, 0.0, row['Grid Size'], row['Block Size']])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  (per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | avg us | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.2f} | {a[2]} | {a[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
