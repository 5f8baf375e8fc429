"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none
--csv --log-file X.csv python bench.py ...`) into a per-kernel table: launches, total and
average duration, share of the captured GPU time.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md
"""
import collections
import csv
import io
import re
import sys


def kernel_key(full):
    m = re.search(r"pf_kernel<(?:knp::)?(\w+(?:<[^>]*>)?)", full)
    if m:
        return m.group(1)
    m = re.search(r"subwarp_kernel<(?:\(int\))?(\d+), (?:knp::)?(\w+)", full)
    if m:
        return f"{m.group(2)} (subwarp {m.group(1)})"
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", full))


def main(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.defaultdict(lambda: [0, 0.0, ""])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        a = agg[kernel_key(r["Kernel Name"])]
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
        a[2] = r["Grid Size"] + " x " + r["Block Size"]
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"{n} launches captured, {total / 1e6:.3f} ms of GPU time (per-launch times are cold-cache and "
          f"serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share | last grid x block |")
    print("|---|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.2f} | {100 * v[1] / total:.1f}% | {v[2]} |")


if __name__ == "__main__":
    main(sys.argv[1])
