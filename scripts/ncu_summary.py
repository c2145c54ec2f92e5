"""Dev tool: text summary of an .ncu-rep (key raw metrics per launch + top stalled SASS instructions)."""
import csv, io, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = re.compile(r"^Kernel Name$|^Grid Size$|^Block Size$|^gpu__time_duration.sum$|^dram__bytes_(read|write).sum$|^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$|"
                  r"^sm__warps_active.avg.pct_of_peak_sustained_active$|^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed$|^smsp__inst_executed.sum$|"
                  r"^sm__issue_active.avg.pct_of_peak_sustained_elapsed$|^launch__registers_per_thread$|^lts__t_sector_hit_rate.pct$|^lts__t_bytes.sum$|"
                  r"^sm__inst_executed_pipe_(alu|fma|lsu|xu|tmem).avg.pct_of_peak_sustained_active$|^smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|barrier|wait|math_pipe_throttle|mio_throttle|not_selected)_per_issue_active.ratio$|"
                  r"^launch__shared_mem_per_block_dynamic$|^smsp__sass_inst_executed_op_tmem_ldt.sum$")
with open(out, "w") as f:
    f.write("# %s\n# source: %s (ncu --set full --clock-control none; cold-cache, serialised launches)\n%s\n\n" % (out.split("/")[-1], rep.split("/")[-1], note))
    for r in rows[2:]:
        for i, k in enumerate(hdr):
            if want.search(k):
                f.write("%-95s %s %s\n" % (k, r[i], units[i]))
        f.write("\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3 and "# Samples" in srows[1]:
        h = srows[1]
        iS, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        data = [r for r in srows[2:] if len(r) > iS and r[iS].isdigit()]
        tot = sum(int(r[iS]) for r in data) or 1
        f.write("top stalled SASS instructions (first profiled launch; samples, share, executed, instruction)\n")
        for r in sorted(data, key=lambda r: -int(r[iS]))[:25]:
            f.write("%7s %5.1f%% %9s  %s\n" % (r[iS], 100.0 * int(r[iS]) / tot, r[iex], r[isrc][:90]))
