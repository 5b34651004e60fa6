"""Summarise an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`) into the few columns
the roofline needs: duration, DRAM bytes read / written, DRAM and tensor-pipe utilisation, L2 hit rate, registers.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full.csv"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct_elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    use = [(hdr.index(m), n) for m, n in COLS if m in hdr]
    w = csv.writer(sys.stdout)
    w.writerow(["id", "kernel"] + [f"{n}[{units[i]}]" if units[i] else n for i, n in use])
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        w.writerow([r[hdr.index("ID")], name] + [r[i] for i, _ in use])


if __name__ == "__main__":
    main(sys.argv[1])
