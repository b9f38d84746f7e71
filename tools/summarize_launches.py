"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per-kernel shares)."""
import collections
import csv
import sys


def main(path, title, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"][:72]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n\nCommand: `{cmd}`\n(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t / 1e3:.3f} | {t / n:.1f} | {t / tot * 100:.1f}% |")
    gs = 100 * sum(t for k, (n, t) in agg.items() if "gemm_kernel" in k) / tot
    print(f"\nGEMM share of the listed launches: {gs:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
