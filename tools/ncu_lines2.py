"""Per source-line stall samples of ONE kernel of an .ncu-rep: ncu_lines2.py REP KERNEL_REGEX [TOP]."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__waves_per_multiprocessor"]
    for r in rows[2:3]:
        for w in want:
            for i, x in enumerate(rows[0]):
                if x == w:
                    print("%-70s %s %s" % (w, r[i], rows[1][i]))
    out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "source", "--csv",
                          "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    hdr, lines = None, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].strip().isdigit():
            si = hdr["# Samples"]
            if len(r) > si and r[si].isdigit() and int(r[si]) > 0:
                why = sorted(((int(r[i]), h[6:]) for h, i in hdr.items()
                              if h.startswith("stall_") and "(" not in h and i < len(r) and r[i].isdigit()
                              and int(r[i]) > 0), reverse=True)[:3]
                lines.append((int(r[si]), int(r[0]), r[1].strip()[:84], r[hdr["Instructions Executed"]], why))
    tot = sum(l[0] for l in lines)
    print("total samples", tot)
    for s, ln, src, ex, why in sorted(lines, reverse=True)[:top]:
        print("%6d %5.1f%% :%-4d exec=%-9s %-84s %s" % (s, 100.0 * s / tot, ln, ex, src, why))


if __name__ == "__main__":
    main()
