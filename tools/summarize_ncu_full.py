"""Summarise an `ncu --set full` report (read with `ncu -i rep --page raw --csv`) into a markdown table."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(rep, title, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    units = rows[1]
    to_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    print(f"# {title}\n\nCommand: `{cmd}` (after the same command exited 0 without ncu)\n")
    print("| # | kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---:|" * len(WANT))
    for i, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")].replace("CUtensorMap_st", "TMap").replace("unsigned short", "u16")[:60]
        vals = []
        for k, _ in WANT:
            v = r[hdr.index(k)] if k in hdr else ""
            try:
                x = float(v.replace(',', ''))
                u = units[hdr.index(k)] if k in hdr else ""
                if k.startswith("dram__bytes"):
                    x *= to_mb.get(u, 1.0)
                if k == "gpu__time_duration.sum":
                    x *= to_us.get(u, 1.0)
                vals.append(f"{x:.1f}")
            except ValueError:
                vals.append(v)
        print(f"| {i} | `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
