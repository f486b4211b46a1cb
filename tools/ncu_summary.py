"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / profiles/ cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__cluster_size", "cluster"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (32 B) from SMs"),
    ("lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "  of which L2 hits"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__inst_executed.sum", "instructions"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
]


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print(("### " if md else "== ") + r[name_i][:100])
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"{'- ' if md else '  '}{label:45s} {r[i]} {units[i]}   [{key}]")
        print()


if __name__ == "__main__":
    main()
