"""Turn an .ncu-rep (ncu --set full) into a small markdown table for profiles/ (run on the build box, no GPU).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_xxx.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in idx]
    print(f"# ncu --set full summary of `{path}`\n")
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        vals = []
        for m, _ in cols:
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.3f}".rstrip("0").rstrip(".") if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip())
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
