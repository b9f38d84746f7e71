#!/usr/bin/env python
"""profiles/gemm_traffic.json from `ncu --set full` captures of bench.py (one per operand dtype): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of the dominant GEMM instantiation (c_fc: ln_2 fold + 1.702 QuickGELU,
M = 100 864, N = 3072, K = 768) and its tensor-pipe utilisation, stamped with the build id of the library that was
profiled (bench.py refuses the file when the id differs from the loaded library's).

    python tools/make_gemm_traffic.py <build_id> f16=<rep> [bf16=<rep>] > profiles/gemm_traffic.json
"""
import csv
import json
import subprocess
import sys


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    return r[0], r[1], r[2:]


def main():
    build_id = sys.argv[1]
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from eoe_b200 import build as b
    # gemm_source_id: identity of the sources the GEMM kernels are compiled from -- only meaningful if the profiled library
    # was built from the sources in the tree (build_id == source_id())
    out = {"build_id": build_id, "gemm_source_id": b.gemm_source_id() if b.source_id() == build_id else None,
           "batch": 512, "by_dtype": {},
           "how": "ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side "
                  "--dtype <dtype>` after the same command exited 0 without ncu (tools/run_profile_r2.sh); per launch"}
    to_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for arg in sys.argv[2:]:
        dt, rep = arg.split("=")
        h, u, rows = rows_of(rep)
        best = None
        for r in rows:
            name = r[h.index("Kernel Name")]
            flat = name.replace("(int)", "").replace("(bool)", "").replace(" ", "")
            if "gemm_kernel<8," not in flat:                  # EOE_EPI_LNFOLD_QUICKGELU_X1702 = 8: the c_fc instantiation
                continue
            best = r
        if best is None:
            continue

        def val(k, scale=None):
            i = h.index(k)
            x = float(best[i].replace(",", ""))
            return x * (to_b.get(u[i], 1.0) if scale == "bytes" else 1.0)
        rd, wr = val("dram__bytes_read.sum", "bytes"), val("dram__bytes_write.sum", "bytes")
        M, N, K = 100864, 3072, 768
        out["by_dtype"][dt] = {
            "kernel": best[h.index("Kernel Name")][:120],
            "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "algorithmic_bytes_per_launch": 2 * (M * K + N * K + M * N),
            "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "time_us_under_ncu": val("gpu__time_duration.sum") / (1e3 if u[h.index("gpu__time_duration.sum")] == "ns" else 1.0),
            "source": rep.split("/")[-1],
            "note": "DRAM traffic below the algorithmic bytes = part of the 620 MB output was still in L2 (126 MB) when the "
                    "kernel ended / A tiles re-read from L2; above = re-reads",
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
