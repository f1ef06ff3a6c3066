"""Compact per-kernel summary of an `ncu --set full` report: for every distinct (kernel, grid) the LAST captured launch —
duration, DRAM bytes and throughput, tensor-pipe activity, issue activity, L1 / L2 hit rates, occupancy, registers.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt       (no GPU needed)"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB"),
    ("dram__bytes_write.sum", "MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("smsp__issue_active.avg.pct", "issue %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__inst_executed.sum", "Minstr"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):          # `ncu --csv --page raw --log-file x.csv` output (lines before the header are ncu chatter)
        text = open(rep).read()
        out = text[text.index('"ID"'):]
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    last = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        key = (r[idx["Kernel Name"]], r[idx.get("launch__grid_size", 0)])
        last[key] = r
    print(f"# {rep}: ncu --set full --clock-control none (one launch per kernel shown: the last one captured)")
    cols = [w for w in WANT if w[0] in idx]
    print(f"{'kernel':58s} " + " ".join(f"{label:>11s}" for _, label in cols))
    for (name, _), r in last.items():
        vals = []
        for metric, label in cols:
            v, u = r[idx[metric]], units[idx[metric]]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                vals.append(f"{v[:11]:>11s}")
                continue
            if label == "us":
                f = f / 1e3 if u in ("nsecond", "ns") else (f * 1e3 if u in ("msecond", "ms") else f)
            if label == "MB":
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                f *= scale
            if label == "Minstr":
                f /= 1e6
            vals.append(f"{f:11.1f}" if label not in ("regs", "grid", "block") else f"{int(f):11d}")
        short = name.replace("(anonymous namespace)::", "").split("(")[0][:58]
        print(f"{short:58s} " + " ".join(vals))


if __name__ == "__main__":
    main()
