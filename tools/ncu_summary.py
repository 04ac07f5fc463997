"""Digest of `ncu -i x.ncu-rep --page raw --csv` exports: one block per kernel launch with the metrics the roofline
discussion in DESIGN.md uses.   python tools/ncu_summary.py a.csv b.csv ..."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA (FP64 tensor) pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (DFMA) pipe active %"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64-pipe warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr_i is None:
        print(f"# {path}: no kernel rows")
        continue
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    for r in rows[hdr_i + 2:]:
        rec = dict(zip(hdr, r))
        unit = dict(zip(hdr, units))
        print(f"# {path}\nKernel  {rec.get('Kernel Name', '?')}")
        t_ns = None
        for k, label in KEYS:
            if k in rec:
                print(f"  {label:42s} {rec[k]:>18s} {unit.get(k, '')}")
                if k == "gpu__time_duration.sum":
                    v = num(rec[k]); u = unit.get(k, "")
                    t_ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1) if v is not None else None
        rd, wr = num(rec.get("dram__bytes_read.sum", "")), num(rec.get("dram__bytes_write.sum", ""))
        if t_ns and rd is not None and wr is not None:
            mult = lambda k: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit.get(k, "byte"), 1)
            b = rd * mult("dram__bytes_read.sum") + wr * mult("dram__bytes_write.sum")
            print(f"  {'DRAM bytes / duration':42s} {b / t_ns:18.1f} GB/s")
        print()
