#!/usr/bin/env python
"""Top warp-stall locations (SASS level) of one launch in an ncu report captured with --set full --import-source on.

    python tools/ncu_hotspots.py report.ncu-rep <launch index> [top N]
"""
import csv
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    print(r[0][1] if len(r[0]) > 1 else r[0])
    h = r[1]
    rows = [x for x in r[2:] if len(x) == len(h)]
    # the CSV repeats every row twice (all samples / not-issued views): keep one
    seen, uniq = set(), []
    for x in rows:
        key = x[h.index("Address")]
        if key not in seen:
            seen.add(key)
            uniq.append(x)
    rows = uniq
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]

    def I(v):
        try:
            return int(v)
        except ValueError:
            return 0
    tot = sum(I(x[isamp]) for x in rows)
    print("total samples", tot)
    agg = {}
    for x in rows:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + I(x[i])
    print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(rows, key=lambda x: -I(x[isamp]))[:top_n]
    for x in sorted(top, key=lambda x: rows.index(x)):
        st = {h[i]: I(x[i]) for i in stall_cols if I(x[i]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{rows.index(x):5d} {x[isamp]:>6} {x[iex]:>8}  {x[isrc][:72]:72s} {st}")


if __name__ == "__main__":
    main()
