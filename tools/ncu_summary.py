"""Markdown table of the key counters of every launch in an ncu `--set full` report.   python tools/ncu_summary.py rep.ncu-rep [more...]"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("smsp__inst_executed.sum", "warp instr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU smem wavefronts %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle")]

for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    names = [r[ik].replace("void ", "").replace("pgrf::", "").split("(")[0] for r in data]
    print(f"\n#### {rep.split('/')[-1]}\n")
    print("| counter | " + " | ".join(f"{n} #{i}" for i, n in enumerate(names)) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            vals.append(v + (" " + units[i] if units[i] not in ("", "%", "ratio") and key.startswith(("gpu__time", "dram__bytes")) else ""))
        print(f"| {label} | " + " | ".join(vals) + " |")
