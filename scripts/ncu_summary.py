"""Summarise an .ncu-rep: headline metrics of each profiled launch + the hottest source lines.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]   (dev tool; output is what gets
committed under profiles/)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for d in data:
    print("== %s" % d[hdr.index("Kernel Name")][:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-70s %s %s" % (w, d[i], units[i]))
    st = [(float(d[i] or 0), h) for i, h in enumerate(hdr) if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$", h)]
    for v, h in sorted(st, reverse=True)[:6]:
        print("   stall %-62s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, out, seen_kernel = None, [], 0
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        pass
    elif len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[7]), int(r[6]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
print("== hottest source lines (all profiled launches), total warp instructions %d" % tot)
for o in sorted(out, reverse=True)[:nl]:
    print("   %5.1f%% inst %11d  samples %6d  %s:%d  %s" % (100.0 * o[0] / tot, o[0], o[1], o[2], o[3], o[4]))
