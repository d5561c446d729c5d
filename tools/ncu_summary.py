#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the MC kernel into the CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/Y.csv "<header comment>" [attempts]

Reads the report with `ncu -i ... --page raw --csv` (last captured launch), keeps the launch shape, DRAM
bytes, hit rates, issue utilisation and the per-issue stall reasons, and -- when the number of attempts of
the captured launch is given -- prints the DRAM bytes and warp instructions per attempt."""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active")


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
    attempts = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[-1]
    keep = sorted((c, u[i], v[i]) for i, c in enumerate(h)
                  if c in KEEP or (c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")))
    val = {c: (unit, x) for c, unit, x in keep}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(float(val[k][1]) * scale[val[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    inst = float(val["smsp__inst_executed.sum"][1])
    with open(out, "w") as f:
        f.write("# " + header + "\n")
        if attempts:
            f.write(f"# {attempts} attempts in the captured launch: {dram / attempts:.1f} DRAM bytes and "
                    f"{inst / attempts:.0f} warp instructions per attempt\n")
        for r in keep:
            f.write(",".join(r) + "\n")
    print(f"kernel {v[h.index('Kernel Name')][:60]}  dram {dram:.4g} B  inst {inst:.4g}"
          + (f"  per attempt: {dram / attempts:.1f} B, {inst / attempts:.0f} inst" if attempts else ""))


if __name__ == "__main__":
    main()
